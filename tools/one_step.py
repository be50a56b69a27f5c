#!/usr/bin/env python
"""Build the BASELINE training-step program and run it eagerly a few times (the command profiled under ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "semantic-segmentation-pytorch_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    from mit_semseg.engine.program import SegProgram
    dev = torch.device("cuda", 0)
    seg = bench.build_model(dev)
    feed = bench.synth_batch(bench.BATCH, bench.CROP, bench.CROP, 8, 304)
    prog = SegProgram(seg, tuple(feed["img_data"].shape), training=True, with_grad=True)
    prog.load_inputs(feed["img_data"].to(dev), feed["seg_label"].to(dev))
    for _ in range(runs):
        prog.run_eager()
    torch.cuda.synchronize()
    print("launches/step", prog.num_launches(), "loss", prog.out[0].item())


if __name__ == "__main__":
    main()
