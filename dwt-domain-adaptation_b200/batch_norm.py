"""Drop-in shim: with ``dwt-domain-adaptation_b200/`` ahead of the reference's ``utils/`` on
sys.path, the reference scripts' ``import batch_norm`` lands here."""
from dwt_b200.batch_norm import *  # noqa: F401,F403
from dwt_b200.batch_norm import _BatchNorm  # noqa: F401
