"""Import hook that runs the reference's Python-2 modules, from where they lie under /root/reference, in this Python 3.

TEST INFRASTRUCTURE: used by tests/golden/make_tf18shim_golden.py only, in the build container (the reference tree does
not exist on the GPU box).  Nothing is copied: a module's source is read, adapted IN MEMORY and executed.  The
adaptations are the mechanical 2to3 ones the hot-path files need, and nothing else:

  * `d.values()[i]` / `d.keys()[i]`  ->  `list(d.values())[i]`   (dict views are not indexable in Python 3)
  * `range(...)` returns a list (the reference concatenates it with lists: `[1, 0] + range(2, n)`)
  * implicit relative imports (`import ed_encoder`, `from ops import map_ta`, `from ed_encoders import ...`) are
    resolved against the reference's package directories
  * the packages' `__init__.py` files are NOT executed (they import every sub-package of nabu, most of it outside the
    hot path and not importable here); packages are empty namespaces

`install(root)` puts the finder on sys.meta_path; `tensorflow` must already resolve to tests/golden/tf18shim.
"""
import builtins
import importlib
import importlib.abc
import importlib.util
import os
import re
import sys

_VIEW = re.compile(r'((?:[A-Za-z_]\w*)(?:\.[A-Za-z_]\w*)*)\.(values|keys|items)\(\)\[')

# where a bare (implicitly relative) module name may live; decoders/ before components/ (both hold a
# beam_search_decoder.py: the bare name is only ever used from inside decoders/)
_BARE_DIRS = ['neuralnetworks/models', 'neuralnetworks/models/ed_encoders', 'neuralnetworks/models/ed_decoders',
              'neuralnetworks/decoders', 'neuralnetworks/components', 'neuralnetworks/trainers']


def _py2_range(*args):
    return list(builtins.range(*args))


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, root, bare):
        self.root = root                      # .../reference (the directory that holds nabu/)
        self.bare = bare                      # the finder of bare names sits LAST on sys.meta_path (real modules win)

    def _locate(self, fullname):
        rel = fullname.replace('.', os.sep)
        path = os.path.join(self.root, rel)
        if os.path.isdir(path):
            return path, True
        if os.path.isfile(path + '.py'):
            return path + '.py', False
        return None, False

    def find_spec(self, fullname, path=None, target=None):
        if (fullname == 'nabu' or fullname.startswith('nabu.')) and not self.bare:
            where, is_pkg = self._locate(fullname)
            if where is None:
                return None
            spec = importlib.util.spec_from_loader(fullname, self, origin=where, is_package=is_pkg)
            if is_pkg:
                spec.submodule_search_locations = [where]
            return spec
        if '.' not in fullname and self.bare:
            for d in _BARE_DIRS:
                canonical = 'nabu.' + d.replace('/', '.') + '.' + fullname
                where, _ = self._locate(canonical)
                if where is not None:
                    return importlib.util.spec_from_loader(fullname, _Alias(canonical), origin=where)
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        where = module.__spec__.origin
        if os.path.isdir(where):
            return                             # a package: empty namespace
        with open(where) as fid:
            src = fid.read()
        src = _VIEW.sub(r'list(\1.\2())[', src)
        module.__file__ = where
        module.__dict__['range'] = _py2_range
        exec(compile(src, where, 'exec'), module.__dict__)


class _Alias(importlib.abc.Loader):
    """a bare name is the canonical dotted module under a second name"""

    def __init__(self, canonical):
        self.canonical = canonical

    def create_module(self, spec):
        return importlib.import_module(self.canonical)

    def exec_module(self, module):
        pass


def install(root='/root/reference'):
    if not os.path.isdir(os.path.join(root, 'nabu')):
        raise RuntimeError('py2ref: no reference tree at %s (this only runs in the build container)' % root)
    sys.meta_path.insert(0, _Finder(root, False))
    sys.meta_path.append(_Finder(root, True))
