from .standard import BasePipeline, StandardPipeline  # noqa: F401
