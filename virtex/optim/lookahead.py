from virtex_b200.optim import Lookahead  # noqa: F401
