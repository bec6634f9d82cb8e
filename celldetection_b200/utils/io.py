"""Model (de)serialisation with the reference's file format (``cd.load_model`` / ``cd.fetch_model`` /
``save_fetchable_model``; /root/reference/celldetection/util/util.py:373-560): a ``torch.save``d dict with keys
``'cd.models'`` (``{model: class name, kwargs, updated_kwargs}``) and ``'state_dict'``."""
import os

import torch


def model2dict(model):
    """util/util.py:527-542"""
    return {'cd.models': dict(model=type(model).__name__, kwargs=dict(getattr(model, 'hparams', {})), updated_kwargs={}),
            'state_dict': model.state_dict()}


def save_fetchable_model(model, filename):
    """util/util.py:545-560 (without the content hash suffix)"""
    torch.save(model2dict(model), filename)
    return filename


def dict2model(conf, **kwargs):
    """util/util.py:373-471: instantiate ``conf['cd.models']['model'](**kwargs)`` and load ``conf['state_dict']``."""
    from .. import models
    meta = conf['cd.models']
    name = meta['model'] if isinstance(meta, dict) else str(meta)
    if not hasattr(models, name):
        raise ValueError(f'{name} is not available in celldetection_b200.models (in scope: {models.__all__})')
    kw = dict(meta.get('kwargs', {})) if isinstance(meta, dict) else {}
    kw.update(meta.get('updated_kwargs', {}) if isinstance(meta, dict) else {})
    kw.update(kwargs)
    # nothing is dropped silently: the constructors accept the options that are identities on this path (pretrained=False,
    # default backbone_kwargs, ...) and raise for every option that would change the network
    model = getattr(models, name)(**kw)
    model.load_state_dict(conf['state_dict'])
    return model.eval()


def load_model(filename, map_location=None, **kwargs):
    """util/util.py:474-479"""
    conf = torch.load(filename, map_location=map_location, weights_only=False)
    if 'cd.models' not in conf:
        raise ValueError('not a celldetection model file (missing "cd.models")')
    return dict2model(conf, **kwargs)


def fetch_model(name, map_location=None, **kwargs):
    """util/util.py:482-509.  Hosted weights need network access (``https://celldetection.org/torch/models/``); this
    build environment has none, so only local files / URLs resolvable by ``torch.hub`` caches work."""
    if os.path.isfile(name):
        return load_model(name, map_location=map_location, **kwargs)
    url = name if name.startswith('http') else f'https://celldetection.org/torch/models/{name}.pt'
    conf = torch.hub.load_state_dict_from_url(url, map_location=map_location, check_hash=kwargs.pop('check_hash', True))
    return dict2model(conf, **kwargs)
