"""Import alias: the package directory is named `gs-livm_b200/` (not a valid Python identifier), so this
module exposes it as `gs_livm_b200` by pointing its __path__ at that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "gs-livm_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
