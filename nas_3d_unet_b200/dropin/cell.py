"""Drop-in shim: put this directory first on sys.path and the reference drivers
(search.py / train.py / prediction.py) import the B200 implementation unchanged."""
from nas_3d_unet_b200.cell import *  # noqa: F401,F403
from nas_3d_unet_b200 import cell as _impl
