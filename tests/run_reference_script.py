"""Run one of the REFERENCE's own scripts (train.py, eval.py, ...; unmodified, read from where they lie) against THIS
package with CUDA replaced by the emulated ABI ("1") or the simulated kernel sources ("sim"):

    python tests/run_reference_script.py <mode> /root/reference/train.py --cfg my.yaml --gpus 0 KEY VALUE ...

`mit_semseg` resolves to semantic-segmentation-pytorch_b200/mit_semseg (sys.path order), so the script's imports - ModelBuilder,
SegmentationModule, TrainDataset, cfg, AverageMeter, parse_devices, user_scattered_collate, ... - are this repository's.
Test infrastructure (tests/test_scripts_support.py); there is no GPU in the build container."""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT, os.path.join(ROOT, "semantic-segmentation-pytorch_b200")):
    sys.path.insert(0, p)


def main():
    mode, script = sys.argv[1], sys.argv[2]
    import conftest
    conftest.install_cuda_stand_in(setattr, mode)
    import mit_semseg
    assert mit_semseg.__file__.startswith(ROOT), mit_semseg.__file__
    import torch
    # DataLoader(pin_memory=True) needs a CUDA context; the stand-in keeps host memory as it is
    real_loader = torch.utils.data.DataLoader

    class Loader(real_loader):
        def __init__(self, *a, **k):
            k["pin_memory"] = False
            super().__init__(*a, **k)
    torch.utils.data.DataLoader = Loader
    if os.environ.get("SSEG_TEST_RANK_CWD") == "1" and "RANK" in os.environ:
        # every rank in its own working directory: a relative cfg.DIR then gives per-rank checkpoint files to compare
        d = os.path.join(os.getcwd(), "rank%s" % os.environ["RANK"])
        os.makedirs(d, exist_ok=True)
        os.chdir(d)
    sys.argv = [script] + sys.argv[3:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
