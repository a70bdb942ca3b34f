from .single_stage import SingleStageDetector  # noqa: F401
