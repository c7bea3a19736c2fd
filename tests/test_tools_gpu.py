"""The reference's tool scripts, byte for byte, on the B200 path: `tools/det/train_codet.py --com disco --kd_flag 1` (one epoch on a
synthetic on-disk dataset in create_data_det.py's format, checkpoint written) and `tools/det/test_codet.py` (validation loop over
the same data with that checkpoint: predict_all -> apply_nms_det -> local mAP bookkeeping), both run with `runpy` after
`disconet_b200.patch.patch_coperception()` -- i.e. exactly `python -m disconet_b200.patch <tool> ...`.
The image lacks the tools' plotting / dataset-SDK imports (matplotlib, nuscenes, mmcv, shapely ...): they are stubbed with mocks
(oracle/ref_import.py), which only the final mean-AP table printing touches."""
import io
import os
import runpy
import sys
import contextlib

import numpy as np
import pytest
import torch

from oracle import ref_import, stage_ref

pytestmark = pytest.mark.gpu


def _run_tool(tool, argv):
    old_argv = sys.argv
    sys.argv = [tool] + argv
    buf = io.StringIO()
    err = None
    try:
        with contextlib.redirect_stdout(buf):
            runpy.run_path(tool, run_name="__main__")
    except SystemExit as e:        # argparse / sys.exit(0)
        err = e if e.code not in (0, None) else None
    except Exception as e:         # noqa: BLE001 -- reported to the caller with the captured output
        err = e
    finally:
        sys.argv = old_argv
    return buf.getvalue(), err


def test_unmodified_train_and_test_codet_tools(cuda_dev, tmp_path):
    if not ref_import.available() or not os.path.exists(stage_ref.tool_path("train_codet.py")):
        pytest.skip("reference package / tools not staged (python -m oracle.stage_ref in the build container)")
    ref_import.install_bypass(mock_heavy=True)
    ref_import.install_stub_shapely()
    import synth_dataset
    from disconet_b200 import DiscoNet, TeacherNet, patch
    A, frames = 2, 2
    root = str(tmp_path)
    torch.manual_seed(1234)          # the tools do not seed: make the two training steps (default init, shuffling) repeatable
    np.random.seed(1234)
    synth_dataset.write_dataset(os.path.join(root, "data"), num_agent=A, n_frames=frames)
    os.makedirs(os.path.join(root, "logs", "disco", "with_rsu"), exist_ok=True)
    patch.patch_coperception()
    try:
        import coperception.models.det as det
        from coperception.configs.Config import Config
        assert det.DiscoNet is DiscoNet and det.TeacherNet is TeacherNet
        teacher = torch.nn.DataParallel(TeacherNet(Config("train", binary=True, only_det=True)))
        torch.save({"epoch": 1, "model_state_dict": teacher.state_dict()}, os.path.join(root, "teacher.pth"))
        common = ["--data", os.path.join(root, "data"), "--com", "disco", "--rsu", "1", "--num_agent", str(A), "--nworker", "0", "--log",
                  "--logpath", os.path.join(root, "logs")]
        out, err = _run_tool(stage_ref.tool_path("train_codet.py"),
                             common + ["--batch_size", "1", "--nepoch", "1", "--auto_resume_path", os.path.join(root, "logs"), "--kd_flag", "1",
                                       "--resume_teacher", os.path.join(root, "teacher.pth")])
        assert err is None, f"train_codet.py failed: {err!r}\n{out[-2000:]}"
        ckpt = os.path.join(root, "logs", "disco", "with_rsu", "epoch_1.pth")
        assert os.path.exists(ckpt), out[-2000:]
        sd = torch.load(ckpt, map_location="cpu")
        assert set(sd) >= {"epoch", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict"} and sd["epoch"] == 1
        assert all(k.startswith("module.") for k in sd["model_state_dict"]) and len(sd["model_state_dict"]) == 321
        assert all(torch.isfinite(v).all() for v in sd["model_state_dict"].values() if v.is_floating_point())
        log = open(os.path.join(root, "logs", "disco", "with_rsu", "log.txt")).read()
        assert "Total loss" in log
        # a model trained for two steps scores either nothing or everything above 0.7: shift the foreground logit so that a few
        # hundred anchors per agent pass the threshold and the NMS path has real work to do (score > 0.7 <=> z1 - z0 > ln(7/3))
        m = DiscoNet(Config("train", binary=True, only_det=True), layer=3, kd_flag=0, num_agent=A)
        m.load_state_dict({k[len("module."):]: v for k, v in sd["model_state_dict"].items()})
        m = m.to(cuda_dev).eval()
        bevs, Ts = [], []
        for a in range(A):
            smp = np.load(os.path.join(root, "data", f"agent{a}", "0_0", "0.npy"), allow_pickle=True).item()
            vox = np.zeros((256, 256, 13), dtype=bool)
            ind = smp["voxel_indices_0"]
            vox[ind[:, 0], ind[:, 1], ind[:, 2]] = 1
            bevs.append(np.rot90(vox, 3).astype(np.float32)[None])
            Ts.append(smp["trans_matrices"])
        with torch.no_grad():
            res, _ = m(torch.from_numpy(np.stack(bevs)).to(cuda_dev), torch.from_numpy(np.stack(Ts))[None], torch.full((1, A), A), batch_size=1)
        margin = (res["cls"][..., 1] - res["cls"][..., 0]).float()            # [A, anchors]
        v = torch.sort(margin.flatten(), descending=True).values
        k = 300
        while k > 1 and float(v[k - 1]) == float(v[k]):                       # never cut inside a run of equal margins
            k -= 1
        shift = float(np.log(7.0 / 3.0)) - 0.5 * (float(v[k - 1]) + float(v[k]))
        assert int((margin + shift > np.log(7.0 / 3.0)).sum(1).max()) <= 2048, "degenerate logits: cannot pick a usable score shift"
        sd["model_state_dict"]["module.classification.conv2.bias"][1::2] += shift
        torch.save(sd, ckpt)
        del m
        out, err = _run_tool(stage_ref.tool_path("test_codet.py"), common + ["--resume", ckpt])
        # the validation loop must have served every frame; what follows it (eval_map's table printing) runs on mocked mmcv /
        # terminaltables and may stop there
        assert out.count("Takes") == frames, f"test_codet.py loop did not finish: {err!r}\n{out[-3000:]}"
        if err is not None:
            import traceback
            tb = "".join(traceback.format_exception(type(err), err, err.__traceback__))
            assert "mean_ap" in tb or "eval_map" in tb, tb[-3000:]
        kept = [int(line.split()[-1]) for line in out.splitlines() if line.startswith("selected:")]
        print("test_codet.py: frames", frames, "| NMS kept per call:", kept[:8], "| tail error:", repr(err)[:120])
    finally:
        patch.unpatch_coperception()
