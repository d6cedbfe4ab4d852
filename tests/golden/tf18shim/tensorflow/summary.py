"""tf.summary: accepted and ignored"""


def scalar(*args, **kwargs):
    return None


image = histogram = scalar
