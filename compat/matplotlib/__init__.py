"""Import-only stand-in for matplotlib (not installed in this image): the reference's utils/vis_utils.py imports it at
module level; the pose / volume plots that would use it are diagnostics and become no-ops."""
import sys
import types


class _Absorb(types.ModuleType):
    """attribute access, calls, context managers and iteration all succeed and do nothing"""

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Absorb(self.__name__ + '.' + name)

    def __call__(self, *a, **k):
        return _Absorb(self.__name__ + '()')

    def __iter__(self):
        return iter(())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __getitem__(self, k):
        return _Absorb(self.__name__ + '[]')


def use(*a, **k):
    return None


for _n in ('pyplot', 'patches', 'cm', 'colors', 'animation'):
    _m = _Absorb(__name__ + '.' + _n)
    sys.modules[__name__ + '.' + _n] = _m
    globals()[_n] = _m
rcParams = {}
