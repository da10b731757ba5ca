#!/usr/bin/env python
"""Result files written by the REFERENCE'S OWN benchmark runner, `tools/test.py`, executed UNMODIFIED on the CPU.

TEST INFRASTRUCTURE ONLY.   python oracle/gen_golden_tool.py
Writes tests/golden/pot_results/<video>.txt (+ fixture.json): what `python tools/test.py --dataset POT210 --config <yaml>
--snapshot <ckpt>` produces on a small synthetic POT-format benchmark (hdn_b200/pot_fixture.py) with the seeded weight fixture.
tests/test_gpu_tool.py runs the GPU twin of that loop (hdn_b200/stream_bench.py) on the same benchmark and compares the files.

The tool's source is read from /root/reference/tools/test.py and executed as __main__ with `__file__` pointing into a scratch
directory, because it resolves its dataset at <tools>/../testing_dataset/POT (tools/test.py:58-62) and /root/reference is
read-only; nothing of it is copied into the repository.  CPU mode = oracle/ref_import.py's shims.
"""
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

ref_import.install()
ref_import.force_homo_backbone_offline()
sys.path.append(REPO)  # behind the reference: only `hdn_b200.*` resolves here, `hdn` / `toolkit` stay the reference's
import torch  # noqa: E402

FIXTURE = {"n_sequences": 2, "n_frames": 10, "size": [360, 480], "seed0": 300, "ext": "png"}
OUT = os.path.join(REPO, "tests", "golden", "pot_results")


def main():
    from hdn_b200 import pot_fixture
    import gen_golden_model as G
    tmp = tempfile.mkdtemp(prefix="hdn_tool_")
    try:
        root = os.path.join(tmp, "testing_dataset", "POT")
        pot_fixture.write_dataset(root, FIXTURE["n_sequences"], FIXTURE["n_frames"], tuple(FIXTURE["size"]), FIXTURE["seed0"], ext=FIXTURE["ext"])
        model, cfg = G.build_reference_model()
        ckpt = os.path.join(tmp, "model", "hdn_fixture.pth")
        os.makedirs(os.path.dirname(ckpt))
        torch.save(model.state_dict(), ckpt)
        del model
        tool = os.path.join(ref_import.REF_ROOT, "tools", "test.py")
        fake = os.path.join(tmp, "tools", "test.py")
        os.makedirs(os.path.dirname(fake))
        argv, cwd = sys.argv, os.getcwd()
        sys.argv = [fake, "--dataset", "POT210", "--config", G.YAML, "--snapshot", ckpt]
        os.chdir(tmp)
        import cv2
        for gui in ("destroyAllWindows", "imshow", "waitKey", "namedWindow"):  # headless OpenCV build: the GUI calls of the tool (:176) are no-ops
            setattr(cv2, gui, lambda *a, **k: None)
        try:
            src = open(tool).read()
            exec(compile(src, fake, "exec"), {"__name__": "__main__", "__file__": fake})
        finally:
            sys.argv = argv
            os.chdir(cwd)
        res = os.path.join(tmp, "results", "POT210", "hdn_fixture")
        shutil.rmtree(OUT, ignore_errors=True)
        os.makedirs(OUT)
        for f in sorted(os.listdir(res)):
            shutil.copy(os.path.join(res, f), os.path.join(OUT, f))
            print(f, sum(1 for _ in open(os.path.join(OUT, f))), "lines")
        with open(os.path.join(OUT, "fixture.json"), "w") as fh:
            json.dump(dict(FIXTURE, tool="tools/test.py --dataset POT210", weights="hdn_b200.synthetic.fill_weights (default calibration)"), fh)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
