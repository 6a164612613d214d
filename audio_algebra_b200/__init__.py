"""Import shim: the package lives in the directory `audio-algebra_b200/` (the name the build
contract fixes); a hyphen is not importable, so `import audio_algebra_b200` resolves here and
executes that directory's __init__.py with this module as the package."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "audio-algebra_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
