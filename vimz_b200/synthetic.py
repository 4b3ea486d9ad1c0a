"""Re-export of vimz_host.synthetic (kept for callers that import it through the product package)."""
from vimz_host.synthetic import *  # noqa: F401,F403
from vimz_host.synthetic import NOVA_AUGMENTED, STEP_CIRCUITS, SyntheticShape  # noqa: F401
