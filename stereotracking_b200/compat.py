"""Minimal stand-ins for the mmengine / mmtrack types at the boundary, used ONLY
when the real packages are not importable (they are not in the build image).

When mmengine + mmtrack are installed the real ``MODELS`` / ``TASK_UTILS``
registries (mmtrack/registry.py:49,:67), ``InstanceData`` and
``TrackDataSample`` (mmtrack/structures/track_data_sample.py:5) are used, so the
classes of ``stereotracking_b200.mmtrack_api`` register into the running
framework under the reference's type names.
"""
import torch

try:  # pragma: no cover - not available in the build container
    from mmengine.structures import InstanceData  # type: ignore
    from mmtrack.registry import MODELS, TASK_UTILS  # type: ignore
    from mmtrack.structures import TrackDataSample  # type: ignore
    HAVE_MMTRACK = True
except Exception:  # ImportError or version asserts of mmtrack/__init__.py:62-73
    HAVE_MMTRACK = False

    class Registry:
        """``@REG.register_module()`` / ``REG.build(dict(type=...))`` subset of
        mmengine.registry.Registry."""

        def __init__(self, name):
            self.name, self.module_dict = name, {}

        def register_module(self, name=None, force=False, module=None):
            def _reg(cls):
                key = name or cls.__name__
                if key in self.module_dict and not force:
                    raise KeyError(f'{key} is already registered in {self.name}')
                self.module_dict[key] = cls
                return cls
            return _reg(module) if module is not None else _reg

        def get(self, key):
            return self.module_dict.get(key.split('.')[-1])

        def build(self, cfg, **default_args):
            cfg = dict(cfg)
            cfg.update({k: v for k, v in default_args.items() if k not in cfg})
            typ = cfg.pop('type')
            cls = self.get(typ) if isinstance(typ, str) else typ
            if cls is None:
                raise KeyError(f'{typ} is not in the {self.name} registry')
            return cls(**cfg)

    MODELS = Registry('model')
    TASK_UTILS = Registry('task util')

    class InstanceData:
        """Field container with attribute / key access (mmengine InstanceData subset)."""

        def __init__(self, **kw):
            object.__setattr__(self, '_fields', {})
            for k, v in kw.items():
                self._fields[k] = v

        def __setattr__(self, k, v):
            self._fields[k] = v

        __setitem__ = __setattr__

        def __getattr__(self, k):
            try:
                return object.__getattribute__(self, '_fields')[k]
            except KeyError:
                raise AttributeError(k)

        def __getitem__(self, item):
            if isinstance(item, str):
                return self._fields[item]
            return InstanceData(**{k: v[item] for k, v in self._fields.items()})

        def __contains__(self, k):
            return k in self._fields

        def get(self, k, default=None):
            return self._fields.get(k, default)

        def keys(self):
            return list(self._fields.keys())

        def clone(self):
            return InstanceData(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in self._fields.items()})

        def __len__(self):
            for v in self._fields.values():
                return len(v)
            return 0

    class TrackDataSample:
        """metainfo + pred_det_instances / pred_track_instances
        (mmtrack/structures/track_data_sample.py:70-92 subset)."""

        def __init__(self, metainfo=None, **fields):
            self._metainfo = dict(metainfo or {})
            self.__dict__.update(fields)

        @property
        def metainfo(self):
            return self._metainfo

        def set_metainfo(self, meta):
            self._metainfo.update(meta)

        @property
        def frame_id(self):
            return self._metainfo.get('frame_id', -1)


def model_shell_base():
    """Base class of the model shells (OCSORT / OCSORT_Disparity).  With mmtrack installed it is the
    reference's own ``BaseMultiObjectTracker(mmengine.model.BaseModel)`` (mmtrack/models/mot/base.py:12,
    :68-112), so a shell built by a Runner has ``data_preprocessor``, ``.to()``, ``.eval()``,
    ``test_step`` and ``forward(mode=...)``.  Without it, a stand-in with the same surface for
    stand-alone use and the tests."""
    if HAVE_MMTRACK:
        from mmtrack.models.mot.base import BaseMultiObjectTracker  # type: ignore
        return BaseMultiObjectTracker

    class BaseMultiObjectTracker:
        """mmengine BaseModel / mmtrack BaseMultiObjectTracker surface used at inference."""

        def __init__(self, data_preprocessor=None, init_cfg=None):
            if isinstance(data_preprocessor, dict):
                data_preprocessor = MODELS.build(data_preprocessor)
            self.data_preprocessor, self.init_cfg = data_preprocessor, init_cfg
            self.training = False

        def to(self, *args, **kwargs):
            return self

        cuda = cpu = to

        def eval(self):
            self.training = False
            return self

        def init_weights(self):
            pass

        def forward(self, inputs, data_samples=None, mode='predict', **kwargs):       # base.py:68-112
            if mode == 'loss':
                return self.loss(inputs, data_samples, **kwargs)
            if mode == 'predict':
                return self.predict(inputs, data_samples, **kwargs)
            if mode == 'tensor':
                return self._forward(inputs, data_samples, **kwargs)
            raise RuntimeError(f'Invalid mode "{mode}". Only supports loss, predict and tensor mode')

        def test_step(self, data):                                                    # mmengine BaseModel.test_step
            if self.data_preprocessor is not None:
                data = self.data_preprocessor(data, False)
            return self.forward(**data, mode='predict') if isinstance(data, dict) else self.forward(*data, mode='predict')

        val_step = test_step

        @property
        def with_detector(self):
            return getattr(self, 'detector', None) is not None

        @property
        def with_motion(self):
            return getattr(self, 'motion', None) is not None

        @property
        def with_tracker(self):
            return getattr(self, 'tracker', None) is not None

    return BaseMultiObjectTracker
