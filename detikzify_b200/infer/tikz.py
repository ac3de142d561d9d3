"""TikzDocument — compile / rasterise a generated TikZ program.

The reference shells out to latexmk + ghostscript + poppler (detikzify/infer/tikz.py:89-156). That is
CPU subprocess work outside the GPU hot path (out of scope) and none of the tools is
installed in the build image, so this class keeps the *interface* the MCTS driver relies on
(``code``, ``is_rasterizable``, ``compiled_with_errors``, ``errors``, ``rasterize()``) and compiles only when
``latexmk`` and ``pdftoppm`` are actually present; otherwise a document is "not rasterizable" and the
search falls back to its compiler-diagnostics reward, exactly as the reference does for failing programs.
A ``backend`` callable can be injected (tests use a deterministic fake renderer).
"""
from __future__ import annotations

import re
import shutil
import subprocess
import tempfile
from functools import cached_property
from os.path import join
from typing import Callable, Dict, Optional

from PIL import Image


class TikzDocument:
    # Callable[[str], Optional[Image.Image]]: code -> rendered image (None = not rasterizable)
    backend: Optional[Callable[[str], Optional[Image.Image]]] = None

    def __init__(self, code: str, timeout: Optional[int] = 60):
        self.code = code
        self.timeout = timeout

    # -- compile -----------------------------------------------------------------------------------
    @cached_property
    def _result(self):
        """(image | None, {line: message})"""
        if TikzDocument.backend is not None:
            try:
                return TikzDocument.backend(self.code), {}
            except Exception as e:  # renderer failure == compile error at an unknown line
                return None, {0: str(e)}
        if not (shutil.which("latexmk") and shutil.which("pdftoppm")):
            return None, {0: "no TeX toolchain available"}
        with tempfile.TemporaryDirectory() as tmp:
            tex = join(tmp, "doc.tex")
            with open(tex, "w") as f:
                f.write(self.code)
            try:
                subprocess.run(["latexmk", "-pdf", "-interaction=nonstopmode", "-halt-on-error", "doc.tex"], cwd=tmp,
                               capture_output=True, timeout=self.timeout)
                subprocess.run(["pdftoppm", "-singlefile", "-png", "-r", "150", "doc.pdf", "doc"], cwd=tmp,
                               capture_output=True, timeout=self.timeout)
                img = Image.open(join(tmp, "doc.png")).convert("RGB")
                img.load()
                errors = self._parse_log(join(tmp, "doc.log"))
                return img, errors
            except Exception:
                return None, self._parse_log(join(tmp, "doc.log")) or {0: "compile failed"}

    @staticmethod
    def _parse_log(path: str) -> Dict[int, str]:
        errors: Dict[int, str] = {}
        try:
            with open(path, errors="replace") as f:
                log = f.read()
        except OSError:
            return errors
        for m in re.finditer(r"^! (.*?)\n(?:.*\n)*?l\.(\d+)", log, flags=re.M):
            errors.setdefault(int(m.group(2)), m.group(1))
        return errors

    # -- interface used by the MCTS driver (reference infer/generate.py:305-343) -----------------------
    @property
    def errors(self) -> Dict[int, str]:
        return self._result[1]

    @property
    def is_rasterizable(self) -> bool:
        return self._result[0] is not None

    @property
    def compiled_with_errors(self) -> bool:
        return bool(self.errors)

    @property
    def has_content(self) -> bool:
        return self.is_rasterizable

    def rasterize(self, size: int = 420) -> Image.Image:
        img = self._result[0]
        if img is None:
            raise ValueError("document is not rasterizable")
        from ..util.image import expand
        return expand(img, size, do_trim=True)

    def save(self, filename: str):
        with open(filename, "w") as f:
            f.write(self.code)
