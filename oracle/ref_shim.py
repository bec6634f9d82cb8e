"""TEST INFRASTRUCTURE ONLY -- import shim for the read-only Python reference.

Makes ``import celldetection`` (from /root/reference) work in a container that lacks
eight of its third-party dependencies (pytorch_lightning, h5py, skimage, ...), so that
the reference itself can be executed on CPU to (a) validate the restatement in
``oracle/cpn_oracle.py`` and (b) mint the golden vectors under ``tests/golden``.

/root/reference does not exist on the GPU box: nothing that runs there may import this
module.  The product package (``celldetection_b200``) never imports anything in
``oracle/``.
"""
import importlib.abc
import importlib.machinery
import inspect
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('CPN_REFERENCE_ROOT', '/root/reference')

_ABSENT = ('pytorch_lightning', 'lightning_fabric', 'h5py', 'skimage', 'matplotlib', 'seaborn', 'timm',
           'segmentation_models_pytorch', 'albumentations', 'imageio', 'tifffile', 'mpi4py', 'lightning')


class _Anything:
    """Permissive placeholder: any attribute / call / subscript yields another placeholder."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and not k and (inspect.isfunction(a[0]) or inspect.isclass(a[0])):
            return a[0]  # behaves like an identity decorator
        return _Anything()

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return _Anything()

    def __getitem__(self, item):
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        full = self.__name__ + '.' + name
        if full in sys.modules:
            return sys.modules[full]
        return _Anything()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in _ABSENT:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def _install_lightning_shims():
    import torch.nn as nn

    class AttributeDict(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v

    class HyperparametersMixin:
        """Just enough of Lightning's mixin: collect the caller's __init__ arguments."""

        def save_hyperparameters(self, *args, ignore=None, frame=None, logger=True):
            frame = frame or inspect.currentframe().f_back
            info = inspect.getargvalues(frame)
            hp = {}
            for name in info.args:
                if name == 'self':
                    continue
                hp[name] = info.locals[name]
            if info.keywords:
                hp.update(info.locals.get(info.keywords, {}))
            self._set_hparams(hp)
            self._hparams_initial = AttributeDict(dict(self._hparams))

        def _set_hparams(self, hp):
            if not hasattr(self, '_hparams') or not isinstance(self.__dict__.get('_hparams'), dict):
                object.__setattr__(self, '_hparams', AttributeDict())
            self._hparams.update(hp)

        @property
        def hparams(self):
            if not hasattr(self, '_hparams'):
                object.__setattr__(self, '_hparams', AttributeDict())
            return self._hparams

        @property
        def hparams_initial(self):
            return getattr(self, '_hparams_initial', AttributeDict())

    class LightningModule(nn.Module, HyperparametersMixin):
        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                import torch
                return torch.device('cpu')

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    pl = sys.modules.get('pytorch_lightning') or importlib.import_module('pytorch_lightning')
    pl.LightningModule = LightningModule
    mixins = importlib.import_module('pytorch_lightning.core.mixins')
    mixins.HyperparametersMixin = HyperparametersMixin
    utilities = importlib.import_module('pytorch_lightning.utilities')
    utilities.rank_zero_only = lambda f: f
    rz = importlib.import_module('pytorch_lightning.utilities.rank_zero')
    rz.rank_zero_only = lambda f: f


_installed = False


def install():
    """Idempotently make ``import celldetection`` resolve to the read-only reference."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f'reference not present at {REFERENCE_ROOT} (expected on the GPU box)')
    sys.meta_path.append(_StubFinder())
    _install_lightning_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    _installed = True


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'celldetection'))


def import_reference():
    install()
    import celldetection as cd
    return cd


class FakeTrainer:
    """Stand-in for ``pl.Trainer.predict`` used by the reference's ``apply_model``."""

    def predict(self, model, dataloaders):
        import torch
        if hasattr(model, 'on_predict_epoch_start'):
            model.on_predict_epoch_start()
        out = []
        with torch.no_grad():
            for i, batch in enumerate(dataloaders):
                out.append(model.predict_step(batch, i))
        return out
