"""Import shim: the product package lives in the directory ``dualvgr-videoqa_b200/`` (a name Python cannot import
directly because of the hyphen). Importing ``dualvgr_videoqa_b200`` executes that directory's ``__init__.py`` as this
module and makes its sub-modules importable as ``dualvgr_videoqa_b200.<name>``."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "dualvgr-videoqa_b200")
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
del _f
