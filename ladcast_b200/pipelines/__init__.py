from .pipeline_AR import AutoRegressive2DPipeline  # noqa: F401
from .scheduler import EDMDPMSolverMultistepScheduler  # noqa: F401
