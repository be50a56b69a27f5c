"""What the reference's scripts need besides the model API so that they run UNMODIFIED on this package: `mit_semseg.config`
(yacs work-alike + the default tree), `mit_semseg.utils`, `mit_semseg.lib.utils.as_numpy`, `mit_semseg.dataset` - and the
proof: the reference's own train.py, read from /root/reference, trains ResNet18dilated+PPM_deepsup on synthetic images for
two epochs through this engine (kernels replaced by the emulated ABI; there is no GPU in the build container), writes the
reference's checkpoint files and resumes from them; the reference's own eval.py then evaluates the checkpoint (multi-scale)
and reports the accuracy / mean IoU the fp32 oracle computes for it."""
import glob
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "mit_semseg")), reason="reference tree not mounted")


def _ref_module(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_config_tree_and_overrides(tmp_path):
    from mit_semseg.config import cfg as global_cfg
    from mit_semseg.config.cfgnode import CfgNode
    cfg = global_cfg.clone()
    assert cfg.MODEL.arch_encoder == "resnet50dilated" and cfg.TRAIN.seed == 304 and cfg.DATASET.imgSizes == (300, 375, 450, 525, 600)
    y = tmp_path / "c.yaml"
    y.write_text('DATASET:\n  imgSizes: (300, 375)\n  padding_constant: 32\nTRAIN:\n  weight_decay: 1e-4\n  lr_encoder: 1\nDIR: "x/y"\n')
    cfg.merge_from_file(str(y))
    assert cfg.DATASET.imgSizes == (300, 375) and cfg.DATASET.padding_constant == 32 and cfg.DIR == "x/y"
    assert cfg.TRAIN.weight_decay == 1e-4 and isinstance(cfg.TRAIN.lr_encoder, float)
    cfg.merge_from_list(["TRAIN.num_epoch", "3", "MODEL.arch_decoder", "c1", "TRAIN.fix_bn", "True"])
    assert cfg.TRAIN.num_epoch == 3 and cfg.MODEL.arch_decoder == "c1" and cfg.TRAIN.fix_bn is True
    cfg.TRAIN.max_iters = 15                      # the scripts add keys on the fly (train.py:264-268)
    assert cfg.TRAIN.max_iters == 15
    with pytest.raises(KeyError):
        cfg.merge_from_list(["TRAIN.no_such_key", "1"])
    with pytest.raises(ValueError):
        cfg.merge_from_list(["TRAIN.num_epoch", "many"])
    if isinstance(cfg, CfgNode):
        text = str(cfg)
        import yaml
        assert yaml.safe_load(text)["TRAIN"]["num_epoch"] == 3      # what train.py writes to config.yaml parses back
        cfg.freeze()
        with pytest.raises(AttributeError):
            cfg.DIR = "z"
        cfg.defrost()
        cfg.DIR = "z"
    assert global_cfg.TRAIN.num_epoch == 20        # the clone was edited, not the global tree


@needs_reference
def test_every_reference_yaml_merges_and_defaults_match_the_reference():
    import ast
    from mit_semseg.config import cfg as mine
    src = open(os.path.join(REF, "mit_semseg/config/defaults.py")).read()
    want = {}
    for node in ast.parse(src).body:         # _C.A.B = literal  (the reference file needs yacs to be imported; parse it instead)
        if isinstance(node, ast.Assign) and isinstance(node.value, (ast.Constant, ast.Tuple, ast.UnaryOp)):
            parts, t = [], node.targets[0]
            while isinstance(t, ast.Attribute):
                parts.append(t.attr)
                t = t.value
            want[".".join(reversed(parts))] = ast.literal_eval(node.value)
    assert len(want) >= 35
    for key, v in want.items():
        node = mine
        for p in key.split("."):
            node = node[p]
        assert node == v and type(node) is type(v), key
    for y in glob.glob(os.path.join(REF, "config", "*.yaml")):
        c = mine.clone()
        c.merge_from_file(y)
        assert isinstance(c.DATASET.imgSizes, tuple) and isinstance(c.TRAIN.weight_decay, float)


@needs_reference
def test_utils_match_the_reference():
    from mit_semseg import utils as U
    from mit_semseg.lib.utils import as_numpy
    R = _ref_module("mit_semseg/utils.py", "_ref_utils")
    rng = np.random.RandomState(0)
    pred, lab = rng.randint(-1, 150, (37, 41)), rng.randint(-1, 150, (37, 41))
    lab[:5] = pred[:5]
    assert U.accuracy(pred, lab)[0] == R.accuracy(pred, lab)[0] and U.accuracy(pred, lab)[1] == R.accuracy(pred, lab)[1]
    for a, b in zip(U.intersectionAndUnion(pred, lab, 150), R.intersectionAndUnion(pred, lab, 150)):
        assert np.array_equal(a, b)
    colors = rng.randint(0, 256, (150, 3)).astype(np.uint8)
    for mode in ("RGB", "BGR"):
        assert np.array_equal(U.colorEncode(lab, colors, mode), R.colorEncode(lab, colors, mode))
    for spec in ("0-3", "0,2", "gpu1-gpu2", "gpu0-2", "3-1", "2,2,gpu2", "gpu5"):
        assert U.parse_devices(spec) == R.parse_devices(spec), spec
    for bad in ("cpu", "1-gpu2", "a-b", ""):
        with pytest.raises(U.NotSupportedCliException):
            U.parse_devices(bad)
        with pytest.raises(R.NotSupportedCliException):
            R.parse_devices(bad)
    m, r = U.AverageMeter(), R.AverageMeter()
    for v, w in ((1.0, 1), (3.0, 2), (0.5, 1)):
        m.update(v, w), r.update(v, w)
    assert m.average() == r.average() and m.value() == r.value() and m.count == r.count
    got = U.unique(lab, return_counts=True)
    assert np.array_equal(got[0], np.unique(lab)) and got[1].sum() == lab.size
    out = as_numpy({"a": torch.ones(2), "b": [torch.zeros(1), 3]})
    assert isinstance(out["a"], np.ndarray) and isinstance(out["b"][0], np.ndarray) and out["b"][1] == 3


def _initial_weights(tmp_path, enc="resnet18dilated", dec="ppm_deepsup", fc=512):
    sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation-pytorch_b200"))
    from mit_semseg.models import models as M, resnet
    torch.manual_seed(1)
    e = M.ResnetDilated(resnet.resnet18(pretrained=False), dilate_scale=8)     # build_encoder(weights='') would download
    d = M.ModelBuilder.build_decoder(arch=dec, fc_dim=fc, num_class=150)
    pe, pd = str(tmp_path / "enc0.pth"), str(tmp_path / "dec0.pth")
    torch.save(e.state_dict(), pe), torch.save(d.state_dict(), pd)
    return pe, pd


@needs_reference
def test_the_reference_train_script_runs_unmodified_on_this_engine(tmp_path):
    from oracle import synth_images as S
    data = tmp_path / "data"
    recs = S.write_dataset(str(data))
    odgt = tmp_path / "train.odgt"
    odgt.write_text("".join(json.dumps(r) + "\n" for r in recs))
    pe, pd = _initial_weights(tmp_path)
    ckpt = tmp_path / "ckpt"
    y = tmp_path / "tiny.yaml"
    y.write_text('DATASET:\n  root_dataset: "%s"\n  list_train: "%s"\n  imgSizes: (64, 80)\n  imgMaxSize: 128\n'
                 'MODEL:\n  arch_encoder: "resnet18dilated"\n  arch_decoder: "ppm_deepsup"\n  fc_dim: 512\n'
                 'TRAIN:\n  batch_size_per_gpu: 2\n  num_epoch: 2\n  epoch_iters: 3\n  workers: 0\n  disp_iter: 1\n'
                 'DIR: "%s"\n' % (data, odgt, ckpt))
    run = [sys.executable, os.path.join(ROOT, "tests", "run_reference_script.py"), "1", os.path.join(REF, "train.py"),
           "--cfg", str(y), "--gpus", "0"]
    out = subprocess.run(run + ["MODEL.weights_encoder", pe, "MODEL.weights_decoder", pd], capture_output=True, text=True,
                         cwd=str(tmp_path), timeout=900)
    log = out.stdout + out.stderr
    assert out.returncode == 0, log[-3000:]
    assert "Training Done!" in log and "Epoch: [2][2/3]" in log
    losses = [float(line.split("Loss: ")[1]) for line in log.splitlines() if "Loss: " in line]
    assert len(losses) == 6 and all(np.isfinite(losses)) and 3.0 < losses[0] < 9.0
    for name in ("encoder_epoch_2.pth", "decoder_epoch_2.pth", "history_epoch_2.pth", "config.yaml"):
        assert (ckpt / name).exists(), name
    hist = torch.load(str(ckpt / "history_epoch_2.pth"))
    assert len(hist["train"]["loss"]) == 6
    # the weights moved, and the files are the reference's format: plain state dicts with the reference's keys
    before, after = torch.load(pe), torch.load(str(ckpt / "encoder_epoch_2.pth"))
    assert list(before) == list(after) and not torch.equal(before["conv1.weight"], after["conv1.weight"])
    # resume (train.py:250-256): start_epoch 2 loads encoder_epoch_2 / decoder_epoch_2 and trains epoch 3
    out = subprocess.run(run + ["TRAIN.start_epoch", "2", "TRAIN.num_epoch", "3"], capture_output=True, text=True,
                         cwd=str(tmp_path), timeout=900)
    log = out.stdout + out.stderr
    assert out.returncode == 0, log[-3000:]
    assert "Loading weights for net_encoder" in log and "Epoch: [3][0/3]" in log and (ckpt / "encoder_epoch_3.pth").exists()

    # ---- the reference's eval.py, unmodified, on the epoch-2 checkpoint (multi-scale, one image at a time); it reads
    # data/color150.mat relative to the working directory, hence cwd = the reference tree (nothing is written there)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_script.py"), "1", os.path.join(REF, "eval.py"),
                          "--cfg", str(y), "--gpu", "0", "DATASET.list_val", str(odgt), "VAL.checkpoint", "epoch_2.pth"],
                         capture_output=True, text=True, cwd=REF, timeout=900)
    log = out.stdout + out.stderr
    assert out.returncode == 0 and "Evaluation Done!" in log, log[-3000:]
    summary = [line for line in log.splitlines() if line.startswith("Mean IoU")][0]
    miou, acc = float(summary.split("Mean IoU: ")[1].split(",")[0]), float(summary.split("Accuracy: ")[1].split("%")[0])
    # the same evaluation with the oracle (the reference's fp32 arithmetic) on the same checkpoint and images
    import copy
    from mit_semseg import dataset as D, utils as U
    from oracle import segnet_oracle as O
    val = D.ValDataset(str(data), copy.deepcopy(recs), S.dataset_options(imgSizes=(64, 80), imgMaxSize=128))
    esd, dsd = torch.load(str(ckpt / "encoder_epoch_2.pth")), torch.load(str(ckpt / "decoder_epoch_2.pth"))
    accm, im, um = U.AverageMeter(), U.AverageMeter(), U.AverageMeter()
    for i in range(len(val)):
        item = val[i]
        lab = item["seg_label"][0].numpy()
        scores = torch.zeros(1, 150, *lab.shape)
        for img in item["img_data"]:
            with torch.no_grad():
                scores = scores + O.segmentation_forward({"img_data": img}, esd, dsd, "resnet18dilated", "ppm_deepsup",
                                                         O.BNState(False), None, segSize=lab.shape) / len(item["img_data"])
        pred = scores.argmax(1)[0].numpy()
        a, pix = U.accuracy(pred, lab)
        inter, union = U.intersectionAndUnion(pred, lab, 150)
        accm.update(a, pix), im.update(inter), um.update(union)
    want_miou, want_acc = float((im.sum / (um.sum + 1e-10)).mean()), accm.average() * 100
    assert abs(acc - want_acc) <= 0.5 and abs(miou - want_miou) <= 1e-3, (summary, want_miou, want_acc)

    # ---- the same eval.py with SSEG_ACCURATE_INFERENCE=1 (fp32-grade inference on bf16 pairs, BASELINE config 2): the printed
    # summary is the fp32 oracle's to the last digit
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_script.py"), "1", os.path.join(REF, "eval.py"),
                          "--cfg", str(y), "--gpu", "0", "DATASET.list_val", str(odgt), "VAL.checkpoint", "epoch_2.pth"],
                         capture_output=True, text=True, cwd=REF, timeout=900, env=dict(os.environ, SSEG_ACCURATE_INFERENCE="1"))
    log = out.stdout + out.stderr
    assert out.returncode == 0, log[-3000:]
    exact = [line for line in log.splitlines() if line.startswith("Mean IoU")][0]
    assert exact.startswith("Mean IoU: {:.4f}, Accuracy: {:.2f}%".format(want_miou, want_acc)), (exact, want_miou, want_acc)

    # ---- eval_multipro.py (one worker process per GPU, results through a queue): same numbers
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_script.py"), "1", os.path.join(REF, "eval_multipro.py"),
                          "--cfg", str(y), "--gpus", "0", "DATASET.list_val", str(odgt), "VAL.checkpoint", "epoch_2.pth"],
                         capture_output=True, text=True, cwd=REF, timeout=900)
    log = out.stdout + out.stderr
    assert out.returncode == 0 and "Evaluation Done!" in log, log[-3000:]
    summary2 = [line for line in log.splitlines() if line.startswith("Mean IoU")][0]
    assert summary2.split(", Inference")[0] == summary.split(", Inference")[0], (summary, summary2)

    # ---- the reference's test.py, unmodified: one image in, class percentages on stdout, a colour rendering on disk
    res = tmp_path / "result"
    res.mkdir()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_script.py"), "1", os.path.join(REF, "test.py"),
                          "--imgs", str(data / "images" / "im_02.png"), "--cfg", str(y), "--gpu", "0",
                          "TEST.checkpoint", "epoch_2.pth", "TEST.result", str(res)],
                         capture_output=True, text=True, cwd=REF, timeout=900)
    log = out.stdout + out.stderr
    assert out.returncode == 0 and "Inference done!" in log and "Predictions in [" in log, log[-3000:]
    from PIL import Image
    vis = Image.open(str(res / "im_02.png"))
    assert vis.size == (2 * 64, 64)          # input image next to the colour-coded prediction


_SIM_TWO_PROCESS = pytest.param("sim", marks=pytest.mark.skipif(os.environ.get("SSEG_TEST_SIM_TWO_PROCESS", "0") != "1",
                                                                reason="1 minute; set SSEG_TEST_SIM_TWO_PROCESS=1 "
                                                                       "(profiles/r1_reference_train_py_two_ranks_on_simulator.log)"))


@needs_reference
@pytest.mark.parametrize("mode", ["1", _SIM_TWO_PROCESS])
def test_the_reference_train_script_as_one_process_per_gpu(tmp_path, mode):
    """`python -m torch.distributed.run --nproc-per-node 2 train.py --gpus 0-1`: the unmodified script knows
    nothing about torch.distributed; wrapping the model in UserScatteredDataParallel joins the job (ensure_process_group),
    every rank consumes its own entry of the loader's per-GPU list, SyncBN statistics and gradients are all-reduced by the
    engine. mode "1": emulated ABI, collectives through the NCCL-mode schedule (gloo). mode "sim": the kernel sources on the
    CPU simulator with the DEFAULT schedule - SyncBN statistics pooled by the csrc/peer.cu kernels through peer arenas that
    the simulator backs with POSIX shared memory, so the two PROCESSES really poll each other's flags. Proof of
    synchronisation: the ranks see different data, each writes its checkpoint into its own directory, the files are identical."""
    import random
    from oracle import synth_images as S
    data = tmp_path / "data"
    recs = S.write_dataset(str(data))
    odgt = tmp_path / "train.odgt"
    odgt.write_text("".join(json.dumps(r) + "\n" for r in recs))
    pe, pd = _initial_weights(tmp_path)
    y = tmp_path / "tiny.yaml"
    y.write_text('DATASET:\n  root_dataset: "%s"\n  list_train: "%s"\n  imgSizes: (64, 80)\n  imgMaxSize: 128\n'
                 'MODEL:\n  arch_encoder: "resnet18dilated"\n  arch_decoder: "ppm_deepsup"\n  fc_dim: 512\n'
                 'TRAIN:\n  batch_size_per_gpu: 2\n  num_epoch: 1\n  epoch_iters: 3\n  workers: 0\n  disp_iter: 1\n'
                 'DIR: "ckpt"\n' % (data, odgt))
    env = dict(os.environ, SSEG_TEST_RANK_CWD="1", SSEG_CKPT_ALL_RANKS="1")   # (by default only rank 0 writes checkpoints)
    if mode == "1":
        env["SSEG_PEER_SYNC"] = "0"
    else:
        env.update(SSEG_DRY_RUN_SMS="16", CUSIM_SMS="16")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(29600 + random.randint(0, 300)), os.path.join(ROOT, "tests", "run_reference_script.py"), mode,
                          os.path.join(REF, "train.py"), "--cfg", str(y), "--gpus", "0-1", "MODEL.weights_encoder", pe,
                          "MODEL.weights_decoder", pd], capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=900)
    log = out.stdout + out.stderr
    assert out.returncode == 0 and log.count("Training Done!") == 2, log[-3000:]
    first = [float(line.split("Loss: ")[1]) for line in log.splitlines() if "Epoch: [1][0/3]" in line]
    assert len(first) == 2 and first[0] != first[1]                  # the ranks really trained on different entries
    for name in ("encoder_epoch_1.pth", "decoder_epoch_1.pth"):
        a, b = torch.load(str(tmp_path / "rank0" / "ckpt" / name)), torch.load(str(tmp_path / "rank1" / "ckpt" / name))
        assert list(a) == list(b)
        for k in a:
            assert torch.equal(a[k], b[k]), (name, k)                 # weights AND BatchNorm running statistics
    start = torch.load(pe)
    assert not torch.equal(start["conv1.weight"], torch.load(str(tmp_path / "rank0" / "ckpt" / "encoder_epoch_1.pth"))["conv1.weight"])
