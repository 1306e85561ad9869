"""Import alias: the package directory is named ``rl-itg_b200`` (not a valid Python identifier), so this
module exposes it as ``import rlitg_b200`` by pointing its ``__path__`` at that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "rl-itg_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
