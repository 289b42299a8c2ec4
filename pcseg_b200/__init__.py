"""Importable alias of the package directory `point-cloud-cnn-segmentation_b200/` (whose name is not
a valid Python identifier).  `import pcseg_b200` resolves sub-modules from that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "point-cloud-cnn-segmentation_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
