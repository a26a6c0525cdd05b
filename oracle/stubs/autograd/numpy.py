from numpy import *  # noqa: F401,F403
