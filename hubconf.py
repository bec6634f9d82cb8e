"""torch.hub entry point mirroring the reference's ``hubconf.py`` (``/root/reference/hubconf.py:16-36``):

    model = torch.hub.load('<this repo>', 'ginoro')

``ginoro`` = CPN + UNet + ResNeXt101 (``CpnResNeXt101UNet``).  Hosted weights need network access."""
dependencies = ['torch']


def ginoro(pretrained: bool = True, pretrained_strict=True, device=None, **kwargs):
    """Ginoro: CPN + UNet + ResNeXt101 (https://proceedings.mlr.press/v212/upschulte23a/upschulte23a.pdf)."""
    import celldetection_b200 as cd
    if not pretrained:
        model = cd.models.CpnResNeXt101UNet(3, **kwargs)
    else:
        model = cd.fetch_model('ginoro_CpnResNeXt101UNet-fbe875f1a3e5ce2c', map_location=device, **kwargs)
    return model.to(device) if device is not None else model
