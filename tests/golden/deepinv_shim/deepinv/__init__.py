"""Minimal stand-in for the `deepinv` package (pinned by the reference at
git+https://github.com/deepinv/deepinv.git@v.0.2.0, requirements.txt:1), which is
neither vendored under /root/reference nor installed in this image.

TEST INFRASTRUCTURE ONLY: it exists so that tests/golden/make_golden.py can import the
reference's *own, unmodified* modules (physics, transforms, losses, crop) and run them
to produce golden vectors.  The definitions below are restated from memory of deepinv
v0.2.0 and are the written specification for everything the reference delegates to
deepinv ("parity unpinned" at this boundary: the reference ships no test that pins
them).  Nothing in the product path imports this package.
"""
from . import physics, loss, transform, models  # noqa: F401
