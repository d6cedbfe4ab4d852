"""tensorflow/python/util/nest.py: flatten / pack_sequence_as / map_structure over tuples, namedtuples, lists, dicts
(sorted keys); anything else - a Tensor, a TensorArray, a python scalar - is a leaf; () flattens to nothing."""


def is_sequence(x):
    return isinstance(x, (tuple, list, dict))


def _children(x):
    if isinstance(x, dict):
        return [x[k] for k in sorted(x)]
    return list(x)


def _rebuild(like, children):
    if isinstance(like, dict):
        return type(like)(zip(sorted(like), children))
    if isinstance(like, tuple) and hasattr(like, '_fields'):
        return type(like)(*children)
    return type(like)(children)


def flatten(x):
    if not is_sequence(x):
        return [x]
    out = []
    for c in _children(x):
        out.extend(flatten(c))
    return out


def pack_sequence_as(structure, flat):
    flat = list(flat)

    def build(s):
        if not is_sequence(s):
            return flat.pop(0)
        return _rebuild(s, [build(c) for c in _children(s)])
    return build(structure)


def map_structure(fn, *structures, **kwargs):
    flats = [flatten(s) for s in structures]
    assert all(len(f) == len(flats[0]) for f in flats), 'nest.map_structure: structures differ'
    return pack_sequence_as(structures[0], [fn(*args) for args in zip(*flats)])


def assert_same_structure(a, b, check_types=True):
    assert len(flatten(a)) == len(flatten(b))
