"""Re-export of vimz_host.field (the buffer helpers live in a package that does not load the shared library)."""
from vimz_host.field import *  # noqa: F401,F403
from vimz_host.field import BN254_P, BN254_R, CURVES, PALLAS_P, R, VESTA_P, CurveInfo  # noqa: F401
