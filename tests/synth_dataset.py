"""Synthetic on-disk samples in the format tools/det/create_data_det.py writes, so that the reference's
UNMODIFIED tools/det/train_codet.py / test_codet.py can be driven end to end without the V2X-Sim dataset (TEST INFRASTRUCTURE).

Layout: <root>/agent{i}/<scene>_<frame>/0.npy = np.save(dict) with the keys V2XSimDet.pick_single_agent / NuscenesDataset read
(datasets/V2XSimDet.py:202-411): sparse voxel indices (per-agent, teacher), sparse labels / regression targets + their masks,
gt boxes, the per-agent transformation matrices, ids."""
import os

import numpy as np

from disconet_b200 import synth


def write_dataset(root: str, num_agent: int = 2, n_frames: int = 2, seed: int = 0, occupancy: float = 0.03) -> str:
    rng = np.random.default_rng(seed)
    for f in range(n_frames):
        T = synth.synth_poses(1, num_agent, seed=seed + 10 * f).numpy()[0]          # T[x, y] = inv(P_x) @ P_y
        teacher = (rng.random((256, 256, 13)) < occupancy)
        for a in range(num_agent):
            d = os.path.join(root, f"agent{a}", f"0_{f}")
            os.makedirs(d, exist_ok=True)
            vox = rng.random((256, 256, 13)) < occupancy
            alloc = rng.random((256, 256, 6)) < 2e-3                                  # anchors assigned to an object
            k = int(alloc.sum())
            reg_mask = np.zeros((256, 256, 6, 1), dtype=bool)
            reg_mask[alloc] = True
            trans = np.zeros((num_agent, 4, 4))
            trans[:num_agent] = T[a]
            sample = {
                "voxel_indices_0": np.argwhere(vox).astype(np.int32),
                "voxel_indices_teacher": np.argwhere(teacher).astype(np.int32),
                "voxel_indices_teacher_no_cross_road": np.argwhere(teacher).astype(np.int32),
                "allocation_mask": alloc,
                "label_sparse": np.ones(k, dtype=np.int8),
                "reg_target_sparse": (rng.standard_normal((k, 1, 6)) * 0.1),
                "reg_loss_mask": reg_mask,
                "gt_max_iou": np.concatenate([np.argwhere(alloc)[: min(k, 20)], np.ones((min(k, 20), 1), dtype=np.int64)], 1),
                "vis_occupy_indices": np.zeros((4, 0), dtype=np.uint8),
                "vis_free_indices": np.zeros((4, 0), dtype=np.uint8),
                "target_agent_id": a,
                "num_sensor": num_agent,
                "trans_matrices": trans,
                "trans_matrices_no_cross_road": trans,
            }
            np.save(os.path.join(d, "0.npy"), sample, allow_pickle=True)
    return root
