"""torchrun script (N >= 1 GPUs): BASELINE config 3 in miniature — one micrograph split over the ranks (contiguous tile bands, rank 0
also the full-image pass), ONE size all-gather + ONE byte-packed payload all-gather of the surviving bit-packed instances, identical
global stages on every rank — must reproduce the golden result of the reference's tile_based_inference_pipeline
(tests/golden/flows_golden.npz, cases "tile_pipeline" and "tile_pipeline_no_edge_filter") on EVERY rank.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_split_micrograph.py"""
import os
import sys

import cv2
import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "golden"))
import flow_cases  # noqa: E402
from deepemia_b200 import batched, distributed as D, engine, synthetic as syn  # noqa: E402
from deepemia_b200.functions import inference as inf  # noqa: E402


def head_batch(pred, images, dev):
    parts = [pred.raw_heads(im) for im in images]
    H, W = images[0].shape[:2]
    in_h, in_w = parts[0][4]
    off = np.concatenate([[0], np.cumsum([len(p[2]) for p in parts])]).astype(np.int64)
    cat = lambda k, dt: torch.as_tensor(np.ascontiguousarray(np.concatenate([p[k] for p in parts]).astype(dt)), device=dev)
    return batched.HeadBatch(cat(0, np.float32), cat(1, np.float32), cat(2, np.float32), cat(3, np.int32), off, H, W,
                             scale_x=float(W) / in_w, scale_y=float(H) / in_h)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    gold = np.load(os.path.join(HERE, "golden", "flows_golden.npz"))
    for name in ("tile_pipeline", "tile_pipeline_no_edge_filter"):
        case = flow_cases.CASES[name]
        image = flow_cases.make_image(case["image_seed"], *case["shape"])
        pred = syn.FakeHeadPredictor(**case["predictors"][0])
        target, small, conf = case["args"]
        kw = case["kwargs"]
        ts, ov, up = kw["tile_size"], kw["overlap_ratio"], kw["upscale_factor"]
        tiles = inf.generate_tiles_with_overlap(image, ts, ov)
        t0, t1 = D.band_of_rank(len(tiles), rank, world)
        ups = [cv2.resize(t, (int(ts * up), int(ts * up)), interpolation=cv2.INTER_LINEAR) for t, _, _ in tiles[t0:t1]]
        full_hb = head_batch(pred, [image], dev) if rank == 0 else None
        tile_hb = head_batch(pred, ups, dev) if ups else None
        xy = np.array([[x, y] for _, x, y in tiles[t0:t1]], np.int32).reshape(-1, 2)
        p = batched.ClassParams(target, target in small, conf, kw.get("iou_threshold", 0.7))
        res = D.split_micrograph(full_hb, tile_hb, xy, image.shape[:2], ts, ov, [p], edge_filter_enabled=kw.get("edge_filter_enabled", True))
        ids = res["per_class"][0]
        n = len(gold[f"{name}/scores"])
        h, w = (int(v) for v in gold[f"{name}/shape"])
        ref = np.unpackbits(gold[f"{name}/bits"], axis=1)[:, :h * w].reshape(n, h, w).astype(bool)
        assert len(ids) == n, (len(ids), n)
        got = engine.unpack_masks(res["iset"], ids).cpu().numpy() != 0
        sc = res["iset"].scores.cpu().numpy()
        for i in range(n):
            assert np.array_equal(got[i], ref[i]), f"rank {rank} {name} mask {i}"
            assert float(sc[ids[i]]) == float(gold[f"{name}/scores"][i])
        print(f"rank {rank}/{world}: {name} OK ({n} instances, {t1 - t0} of {len(tiles)} tiles on this rank)", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
