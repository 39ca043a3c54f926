from . import tools, models  # noqa: F401
