"""Import shim: the package directory is literally ``pdesolver.jl_b200/`` (the
layout the build contract names), which is not an importable identifier, so
this module turns itself into that package under the name ``pdesolver_jl_b200``.
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "pdesolver.jl_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
__package__ = __name__
__spec__.submodule_search_locations = __path__
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
