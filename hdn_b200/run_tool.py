"""Run one of the reference's tools UNCHANGED against the B200 packages.

    python -m hdn_b200.run_tool <reference checkout>/tools/test.py --dataset POT210 --config <yaml> --snapshot <ckpt>
    python -m hdn_b200.run_tool <reference checkout>/tools/demo.py --config <yaml> --snapshot <ckpt> --video <file>

The tool file is executed as __main__ exactly as `python tools/test.py ...` would, with one difference: the directory of
the mirrored packages (hdn_b200/compat: `hdn`, `homo_estimator`, `toolkit`) is put in front of sys.path, so every
`from hdn... import` of the tool (tools/test.py:14-19, tools/demo.py:15-19) resolves to the sm_100a implementation instead of
the reference's own packages.  Nothing in the tool is edited or copied.  On a headless OpenCV build the tool's GUI calls
(cv2.destroyAllWindows at tools/test.py:176, imshow / waitKey under --vis) are replaced by no-ops.
"""
import os
import runpy
import sys


def activate_for_tool(tool_path):
    from hdn_b200 import compat
    compat_dir = compat.activate()
    # `python tools/test.py` would put the tool's directory first; keep it, but behind the mirror packages
    tool_dir = os.path.dirname(os.path.abspath(tool_path))
    sys.path[:] = [compat_dir] + [p for p in sys.path if p != compat_dir]
    if tool_dir not in sys.path:
        sys.path.insert(1, tool_dir)
    import cv2
    try:
        cv2.destroyAllWindows()
    except cv2.error:  # no GUI backend in this build
        for gui in ("destroyAllWindows", "imshow", "waitKey", "namedWindow", "setMouseCallback"):
            setattr(cv2, gui, lambda *a, **k: -1)
    return compat_dir


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 2
    tool = argv[0]
    if not os.path.isfile(tool):
        raise SystemExit("run_tool: %s is not a file" % tool)
    activate_for_tool(tool)
    sys.argv = [tool] + argv[1:]
    runpy.run_path(tool, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
