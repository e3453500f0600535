"""Launcher for UNMODIFIED reference entry scripts:  ``python -m seedx_b200.run src/inference/eval_img2text_seed_x_i.py [args]``.

The reference's scripts start with ``import hydra`` / ``import pyrootutils`` / ``from omegaconf import OmegaConf`` /
``from diffusers import …`` (/root/reference/src/inference/eval_img2text_seed_x_i.py:1-14) and resolve ``src.…`` `_target_`s through the
``.project-root`` marker.  This launcher registers the stand-ins for whichever of those third-party packages are missing
(seedx_b200.compat.install), then executes the script as ``__main__`` with the script's directory first on ``sys.path`` — exactly what
``python3 <script>`` does — so a script copied from the reference runs against this repo's ``src/`` and ``configs/`` without edits.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m seedx_b200.run <script.py> [script args]")
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        raise SystemExit(f"seedx_b200.run: no such script: {argv[0]}")
    from . import compat
    compat.install()
    sys.argv = [script] + argv[1:]
    sys.path.insert(0, os.path.dirname(script))       # `from any_res import process_anyres_image` (sibling import of the scripts)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
