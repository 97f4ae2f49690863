"""Import shim: the package directory is `hw-radiosity_b200/` (not a valid Python identifier);
this module makes it importable as `hw_radiosity_b200`."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "hw-radiosity_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
