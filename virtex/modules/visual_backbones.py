from virtex_b200.modules import VisualBackbone, TorchvisionVisualBackbone  # noqa: F401
