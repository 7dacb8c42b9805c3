import sys
from matplotlib import _Absorb

Axes3D = _Absorb('mpl_toolkits.mplot3d.Axes3D')
art3d = _Absorb('mpl_toolkits.mplot3d.art3d')
sys.modules[__name__ + '.art3d'] = art3d
