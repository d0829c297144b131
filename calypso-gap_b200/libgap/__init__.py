"""libgap -- drop-in for the reference's ``libgap`` Python package
(gappy/libgap/__init__.py): ``libgap.libgap`` is the f2py extension module
(fgap_read, fgap_calc, fget_bond, car2acsf, write_array_2dim), ``libgap.GAP`` and
``libgap.BOND`` the thin classes on top of it.  The arithmetic runs on a B200
through lib/libgapcu.so; there is no CPU fallback.
"""
__version__ = "0.1.0"
