"""Import alias: the product package lives in the directory ``gan-aug-pfa_b200/`` (the name the
build contract asks for), which is not a valid Python identifier.  This stub makes it importable as
``gan_aug_pfa_b200`` by pointing the package search path at that directory and executing its
``__init__.py`` in this module's namespace."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gan-aug-pfa_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py"), "r") as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
