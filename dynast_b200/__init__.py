"""Import shim: the package sources live in ``dynast-release_b200/`` (a directory name Python cannot
import directly); this module makes them importable as ``dynast_b200``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'dynast-release_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
