from .cpn import CPN, CpnU22, CpnResNet18FPN, CpnResNeXt101UNet

__all__ = ['CPN', 'CpnU22', 'CpnResNet18FPN', 'CpnResNeXt101UNet']
