"""Importable alias of the ``audio-raytracer_b200/`` source directory.

The product directory carries the name the project layout prescribes
(``audio-raytracer_b200``), which is not a valid Python identifier; this shim
package extends its ``__path__`` there so ``import audio_raytracer_b200`` works.
"""
import os as _os

_src = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "audio-raytracer_b200")
__path__.insert(0, _src)

with open(_os.path.join(_src, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_src, "__init__.py"), "exec"))
del _f
