from .th import *  # noqa: F401,F403
