"""Import shim: the package directory is `speech-inpainting_b200/` (repo layout contract), which
is not a valid Python identifier; this module exposes it as `speech_inpainting_b200`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "speech-inpainting_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
