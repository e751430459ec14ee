"""Import shim: the package directory is `lifelong-clip_b200/` (a hyphen is not importable), so this
module gives it the importable name `lifelong_clip_b200` by pointing __path__ at that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "lifelong-clip_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
