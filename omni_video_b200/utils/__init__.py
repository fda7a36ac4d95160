from .fm_solvers_unipc import FlowUniPCMultistepScheduler  # noqa: F401
