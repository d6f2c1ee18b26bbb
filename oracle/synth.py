"""Re-export of the seeded synthetic data generators (they live in the package so that bench.py's product arm does not
import anything from oracle/)."""
from avi_talking_b200.synth import *  # noqa: F401,F403
from avi_talking_b200.synth import W2V, N_VERT, N_FACE, N_JOINT, _t  # noqa: F401
