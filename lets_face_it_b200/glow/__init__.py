"""Drop-in for the reference's `glow_pytorch.glow` package (same public names)."""
from .models import FeatureEncoder, FlowNet, FlowStep, Glow, ModalityEncoder, SeqGlow, f_seq  # noqa: F401
from .modules import ActNorm2d, GaussianDiag, InvertibleConv1x1, LinearZeros  # noqa: F401
from .utils import calc_jerk, get_longest_history  # noqa: F401
