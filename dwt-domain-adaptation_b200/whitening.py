"""Drop-in shim: with ``dwt-domain-adaptation_b200/`` ahead of the reference's ``utils/`` on
sys.path, the reference scripts' ``import whitening`` lands here."""
from dwt_b200.whitening import *  # noqa: F401,F403
from dwt_b200.whitening import _Whitening  # noqa: F401
