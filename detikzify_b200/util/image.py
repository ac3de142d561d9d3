"""Host-side image helpers with the reference's semantics (detikzify/util/image.py:24-60):
``load`` (path / bytes / base64 / PIL, alpha composited on white), ``trim`` to content,
``expand`` = optional trim + LANCZOS pad to a square. (URL loading needs network: not supported.)"""
from __future__ import annotations

from base64 import b64decode
from io import BytesIO
from os.path import isfile

from PIL import Image, ImageChops, ImageOps

DUMMY_IMAGE = Image.new("RGB", (24, 24), color="white")


def remove_alpha(image: Image.Image, bg="white") -> Image.Image:
    background = Image.new("RGBA", image.size, bg)
    return Image.alpha_composite(background, image.convert("RGBA")).convert("RGB")


def trim(image: Image.Image, bg="white") -> Image.Image:
    diff = ImageChops.difference(image, Image.new(image.mode, image.size, bg))
    bbox = diff.getbbox()
    return image.crop(bbox) if bbox else image


def expand(image: Image.Image, size: int, do_trim: bool = False, bg="white") -> Image.Image:
    if do_trim:
        image = trim(image, bg=bg)
    return ImageOps.pad(image, (size, size), color=bg, method=Image.Resampling.LANCZOS)


def load(image, bg="white", timeout=None) -> Image.Image:
    if isinstance(image, bytes):
        image = Image.open(BytesIO(image))
    elif isinstance(image, str):
        if isfile(image):
            image = Image.open(image)
        elif image.startswith(("http://", "https://")):
            raise ValueError("remote images are not supported offline")
        else:
            try:
                image = Image.open(BytesIO(b64decode(image.split(",")[-1])))
            except Exception as e:
                raise ValueError(f"Incorrect image source (path, bytes or base64 expected): {e}")
    image = ImageOps.exif_transpose(image)
    return remove_alpha(image, bg=bg)
