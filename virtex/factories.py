from virtex_b200.factories import *  # noqa: F401,F403
from virtex_b200.factories import (Factory, VisualBackboneFactory, TextualHeadFactory, PretrainingModelFactory,  # noqa: F401
                                   CaptionDecoderFactory, OptimizerFactory, LRSchedulerFactory)
