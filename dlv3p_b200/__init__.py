"""Importable alias of the package directory `tf-keras-deeplabv3p-model-set_b200/` (hyphens are not valid
in a Python module name): this package's __path__ points there, so `import dlv3p_b200.ffi` works."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'tf-keras-deeplabv3p-model-set_b200')
__path__.insert(0, _PKG_DIR)

with open(_os.path.join(_PKG_DIR, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, '__init__.py'), 'exec'))
