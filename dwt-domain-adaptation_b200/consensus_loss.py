"""Drop-in shim: with ``dwt-domain-adaptation_b200/`` ahead of the reference's ``utils/`` on
sys.path, the reference scripts' ``import consensus_loss`` lands here."""
from dwt_b200.consensus_loss import *  # noqa: F401,F403
