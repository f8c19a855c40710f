"""Import shim: ``import mobody_b200`` loads the package that physically lives in
``mobody-model-based-off-dynamics-offline-reinforcement-learning_b200/`` (a directory name Python
cannot import directly).  One module identity: everything is registered as ``mobody_b200.*``."""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "mobody-model-based-off-dynamics-offline-reinforcement-learning_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
