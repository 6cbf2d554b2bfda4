"""B200-native depth-guided tracking path for StereoTracking."""
__version__ = '0.1.0'
