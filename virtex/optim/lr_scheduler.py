from virtex_b200.optim import (LinearWarmupNoDecayLR, LinearWarmupMultiStepLR, LinearWarmupLinearDecayLR,  # noqa: F401
                               LinearWarmupCosineAnnealingLR)
