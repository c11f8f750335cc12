"""Import alias for the product package.

The product lives in `critic-guided-segmentation-of-rewarding-objects-in-first-person-views_b200/`
(the name the build contract asks for); hyphens are not importable, so this shim
puts that directory on this package's search path: `import cgs_b200.nets` resolves to
`critic-guided-..._b200/nets.py`.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "critic-guided-segmentation-of-rewarding-objects-in-first-person-views_b200")
__path__.insert(0, _PKG_DIR)
PKG_DIR = _PKG_DIR
