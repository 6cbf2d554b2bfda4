"""Opt-in registration of the model shells.

``custom_imports = dict(imports=['stereotracking_b200.mmtrack_shell'])`` replaces the reference's
``OCSORT`` / ``OCSORT_Disparity`` (mmtrack/models/mot/ocsort.py:12, ocsort_disparity.py:15) with the
shells of ``mmtrack_api`` - ``BaseMultiObjectTracker`` subclasses like the originals, so a Runner's
``model.to()``, ``test_step`` and data preprocessor work unchanged.  Importing
``stereotracking_b200.mmtrack_api`` alone swaps only the tracker and the Kalman filter.
"""
from .mmtrack_api import register_model_shells

OCSORT, OCSORT_Disparity = register_model_shells(force=True)
