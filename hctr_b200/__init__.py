"""Import alias: `hctr_b200` is the package that lives in `handwritten-chinese-ocr-samples_b200/`
(a directory name Python cannot import directly)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "handwritten-chinese-ocr-samples_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
