from .LaDCast_3D_model import LaDCastTransformer3DModel  # noqa: F401
from .DCAE import AutoencoderDC  # noqa: F401,E402
