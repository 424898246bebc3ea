"""Import shim: the package sources live in `glue-factory_b200/` (the name the
build contract fixes; a hyphen is not importable), this module makes them
importable as `gluefactory_b200`."""
import os as _os

_src = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "glue-factory_b200")
__path__.insert(0, _src)
__version__ = "0.1.0"
