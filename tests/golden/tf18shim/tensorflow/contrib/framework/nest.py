from tensorflow.nest_impl import *  # noqa: F401,F403
