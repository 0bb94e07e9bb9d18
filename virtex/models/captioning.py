from virtex_b200.models import *  # noqa: F401,F403
from virtex_b200.models import CaptioningModel, ForwardCaptioningModel, BidirectionalCaptioningModel, VirTexModel  # noqa: F401
