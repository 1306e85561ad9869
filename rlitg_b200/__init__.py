"""rlitg_b200 -- the importable name of the package that lives in ``rl-itg_b200/``.

The repository layout fixes the package directory as ``rl-itg_b200`` (not a Python identifier).  This regular package
extends its ``__path__`` with that directory, so ``import rlitg_b200.kinematics`` resolves to
``rl-itg_b200/kinematics.py`` through the normal import machinery (relative imports inside the package included)."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "rl-itg_b200"))
__version__ = "0.2.0"
