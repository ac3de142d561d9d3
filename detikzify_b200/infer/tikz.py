"""TikzDocument — compile / rasterise a generated TikZ program (the MCTS reward's input).

CPU subprocess work outside the GPU hot path, but its *semantics* decide the MCTS reward, so they
follow the reference (detikzify/infer/tikz.py:38-156):

  * ``latexmk -f -nobibtex -norc -file-line-error -interaction=nonstopmode -<engine>`` — recoverable errors still yield a
    PDF (``-f``, nonstopmode): such a document IS rasterizable and merely ``compiled_with_errors`` (what ``strict`` and the
    diagnostics reward look at); engines pdflatex, lualatex, xelatex are tried in turn and the attempt whose first error
    comes latest is kept (:114-133);
  * ``status`` is the compiler's exit status, ``compiled_with_errors = status != 0`` (:50-52);
  * ``errors`` = ``file:line:error`` entries of the kept log, keyed by line for the root file and 0 for anything else (:54-77);
  * page numbers are suppressed (``\\pagestyle{empty}`` injected after the first line, :96-97) and the LAST page is kept (:104-110);
  * ``rasterize()`` returns ``None`` when there is no image (:139-146).

Not available offline and therefore substituted: pymupdf / pdfCropMargins / pdf2image (page selection, vector crop,
rasterisation) — ``pdftoppm`` renders the last page and the white margins are trimmed on the raster (``util.image.expand``
with ``do_trim``), which is what the crop achieves before scoring. Without ``latexmk`` and ``pdftoppm`` on the PATH a
document is "not rasterizable" and the search falls back to its compiler-diagnostics reward. A ``backend`` callable can be
injected (tests use a deterministic fake renderer).
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import tempfile
from collections import namedtuple
from functools import cached_property
from os.path import isfile, join
from typing import Callable, Dict, List, Optional, Union

from PIL import Image

Output = namedtuple("Output", ["image", "status", "log", "rootfile"], defaults=[None, -1, "", None])


class TikzDocument:
    engines: List[str] = ["pdflatex", "lualatex", "xelatex"]
    # Callable[[str], Optional[Image.Image]]: code -> rendered image (None = not rasterizable)
    backend: Optional[Callable[[str], Optional[Image.Image]]] = None

    def __init__(self, code: str, timeout: Optional[int] = 60):
        self.code = code
        self.timeout = timeout

    @classmethod
    def set_engines(cls, engines: Union[str, list]):
        cls.engines = [engines] if isinstance(engines, str) else list(engines)

    # -- compile -----------------------------------------------------------------------------------
    @cached_property
    def _result(self) -> Output:
        if TikzDocument.backend is not None:
            try:
                img = TikzDocument.backend(self.code)
                return Output(image=img, status=0 if img is not None else 1, log="")
            except Exception as e:  # renderer failure == compile error at an unknown line
                return Output(image=None, status=1, log=f"backend:0:{e}", rootfile="document")
        if not (shutil.which("latexmk") and shutil.which("pdftoppm")):
            return Output(image=None, status=-1, log="", rootfile=None)
        return self.compile()

    def compile(self) -> Output:
        lines = self.code.split("\n")
        cmd = r"\thispagestyle{empty}\pagestyle{empty}"
        lines.insert(1, cmd + r"\AtBeginDocument{" + cmd + "}")      # no page numbers in the rendered page
        best = dict(status=-1, log="", errorln=-1, pdf=None)
        with tempfile.TemporaryDirectory() as tmp:
            root = join(tmp, "tikz.tex")
            with open(root, "w") as f:
                f.write("\n".join(lines))
            open(join(tmp, "tikz.bbl"), "a").close()                   # some classes expect a bibliography file
            kept = join(tmp, "kept.pdf")
            for engine in self.engines:
                try:
                    proc = subprocess.run(
                        ["latexmk", "-f", "-nobibtex", "-norc", "-file-line-error", "-interaction=nonstopmode", f"-{engine}", root],
                        cwd=tmp, timeout=self.timeout, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                        env={**os.environ, "max_print_line": "1000"})
                    status, log = proc.returncode, proc.stdout.decode(errors="ignore")
                except subprocess.TimeoutExpired as e:
                    status, log = -1, (e.output or b"").decode(errors="ignore")
                except FileNotFoundError:
                    break
                if status == 0:
                    best.update(status=0, log="", errorln=1 << 30)
                    if isfile(join(tmp, "tikz.pdf")):
                        shutil.copyfile(join(tmp, "tikz.pdf"), kept)
                    break
                first = re.search(rf"^{re.escape(root)}:(\d+):.+$", log, re.M)
                errorln = int(first.group(1)) if first else 0
                if errorln > best["errorln"]:                          # keep the engine that got furthest
                    best.update(status=status, log=log, errorln=errorln)
                    if isfile(join(tmp, "tikz.pdf")):
                        shutil.copyfile(join(tmp, "tikz.pdf"), kept)
            image = self._render_last_page(kept, tmp) if isfile(kept) else None
        return Output(image=image, status=best["status"], log=best["log"], rootfile=root)

    def _render_last_page(self, pdf: str, tmp: str) -> Optional[Image.Image]:
        try:
            info = subprocess.run(["pdfinfo", pdf], capture_output=True, timeout=self.timeout).stdout.decode(errors="ignore")
            m = re.search(r"^Pages:\s+(\d+)", info, re.M)
            last = m.group(1) if m else "1"
            subprocess.run(["pdftoppm", "-f", last, "-l", last, "-singlefile", "-png", "-r", "150", pdf, join(tmp, "page")],
                           capture_output=True, timeout=self.timeout)
            img = Image.open(join(tmp, "page.png")).convert("RGB")
            img.load()
            return img
        except Exception:
            return None

    # -- interface used by the MCTS driver (reference infer/generate.py:305-343) -----------------------
    @property
    def status(self) -> int:
        return self._result.status

    @property
    def log(self) -> str:
        return self._result.log

    @property
    def compiled_with_errors(self) -> bool:
        return self.status != 0

    @property
    def errors(self) -> Dict[int, str]:
        """{line: message}; line 0 collects errors without a line number in the root file."""
        if not self.compiled_with_errors:
            return {}
        errors: Dict[int, str] = {}
        root = self._result.rootfile
        for file, line, msg in re.findall(r"^(.+):(\d+):(.+)$", self.log, re.M):
            if root is None or file == root:
                errors[int(line)] = msg.strip()
            else:
                errors[0] = msg.strip()
        return errors or {0: "Fatal error occurred, no output PDF file produced!"}

    @property
    def is_rasterizable(self) -> bool:
        return self._result.image is not None

    @property
    def has_content(self) -> bool:
        img = self.rasterize()
        return img is not None and img.getcolors(1) is None

    def rasterize(self, size: int = 420, expand_to_square: bool = True) -> Optional[Image.Image]:
        img = self._result.image
        if img is None:
            return None
        from ..util.image import expand
        if expand_to_square:
            return expand(img, size, do_trim=True)
        return img

    def save(self, filename: str):
        if filename.endswith(".tex"):
            with open(filename, "w") as f:
                f.write(self.code)
        elif (img := self.rasterize()) is not None:
            img.save(filename)
        else:
            raise ValueError(f"Couldn't save {filename!r}: the document has no image")
