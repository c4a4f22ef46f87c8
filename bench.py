#!/usr/bin/env python
"""bench.py -- images/sec of the Mask-RCNN hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload pipeline|custom_layers|roialign] [--batch B] [--config A|small|single]

--workload roialign is BASELINE.json configs[4] (ROIAlign-only microbench, P2 = 256x256x256, level-2 rois; sweep of
R x batch x pool x layout, L2 flushed): metric = ROIAlign HBM GB/s, see run_roialign / tools/bench_roialign.py.

One "step" = one pass of the hot path over one batch of synthetic 1024x1024
inputs (BASELINE.json configs[1]: batch 8 per GPU, 6000 -> 1000 proposals).
  value  : whole-job images/s, inputs already resident in HBM (device pointers
           through the C ABI), CUDA-event timed, max over ranks.
  e2e    : the same through the reference-facing call with HOST buffers (pinned),
           host<->device copies inside the timed region.
  roofline: the dominant kernel class, timed live with CUDA event pairs on the
           library's stream (mrcnn_profile_*), against MEASURED_PEAKS.json.
  cpu_baseline: the oracle (a port; the Swift reference cannot run on Linux) on a
           bounded sample, rank 0 / N=1 only.
--impl reference times that CPU path alone (all host threads it can use).
Nothing here reads /root/reference.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec (1024x1024, 1000 ROIs)"
UNIT = "images/s"
IMG = 1024
# BASELINE.json configs the pipeline workload can be run on: name -> (architecture, image size, proposals, batch or None)
CONFIGS = {"A": (101, 1024, 1000, None),         # configs[1] (batch 8 per GPU) / configs[3] (8 GPUs): the headline
           "small": (50, 512, 300, None),        # configs[2]: ResNet50, 512x512, 300 proposals
           "single": (101, 1024, 1000, 1)}       # configs[0]: one 1024x1024 image per call (latency)


def config_metric(name):
    arch, size, props, _ = CONFIGS[name]
    return METRIC if name != "small" else f"images/sec ({size}x{size}, {props} ROIs)"


# --------------------------------------------------------------------------- utils
def load_traffic(kernel_class):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch for a kernel class, from the committed ncu capture."""
    for name in ("traffic_r2.json", "traffic_r1.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))[kernel_class]["traffic_bytes_per_launch"]
        except (OSError, KeyError, ValueError):
            continue
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "bf16": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------- CPU legs (oracle)
def cpu_custom_layers_image(orc, synth, anchors, image_index, maps):
    """The reference's custom-layer path for ONE image on the CPU (oracle port): full argsort of
    261,888 scores, sequential greedy NMS with double IoU, scalar crop-and-resize."""
    probs, deltas = synth.rpn_outputs(anchors, image_index)
    t0 = time.perf_counter()
    rois, _, _ = orc.proposal(probs, deltas, anchors)
    pooled, _ = orc.pyramid_roialign(rois, maps, 7)
    pr, bb = synth.classifier_outputs(1000, image_index)
    cls = orc.classifier_select(pr, bb)
    det, _, _ = orc.detection(rois, cls)
    pooled14, _ = orc.pyramid_roialign(det, maps, 14)
    dt = time.perf_counter() - t0
    return dt


_CPU_STATE = {}


def cpu_pipeline_image(m, orc, image_index, architecture=101, size=IMG, proposals=1000):
    """The reference's whole path for ONE image on the host CPU: dense graphs = oracle/dense_ref.py (PyTorch CPU
    fp32, all host threads; what Core ML's CPU path does with the .mlmodel graphs), custom layers = oracle.c
    (single thread, like the reference's Swift layers).  Returns (seconds, per-stage seconds)."""
    import torch
    from oracle.dense_ref import Ref
    st = _CPU_STATE
    if st.get("key") != (architecture, size):
        st["key"] = (architecture, size)
        folded, _ = m.weights.synthetic_blobs(architecture)
        st["ref"] = Ref(folded, architecture, act_half=False, device="cpu")
        st["anchors"] = m.synth.generate_anchors(size, size)
    ref, anchors = st["ref"], st["anchors"]
    rng = np.random.default_rng(20260 + image_index)
    img = rng.integers(0, 256, (1, size, size, 3), dtype=np.uint8)
    t = [time.perf_counter()]
    with torch.no_grad():
        fm, probs, deltas = ref.backbone(img)
        t.append(time.perf_counter())
        rois, _, _ = orc.proposal(probs[0].numpy(), deltas[0].numpy(), anchors, pre_nms=6000, max_proposals=proposals)
        t.append(time.perf_counter())
        maps = [np.ascontiguousarray(f[0].permute(2, 0, 1).numpy()) for f in fm]      # CHW fp32, the layer's layout
        pooled, _ = orc.pyramid_roialign(rois, maps, 7, size, size)
        t.append(time.perf_counter())
        pr, bb, _ = ref.classifier(np.ascontiguousarray(pooled.transpose(0, 2, 3, 1)))
        cls = orc.classifier_select(pr.numpy(), bb.numpy())
        t.append(time.perf_counter())
        det, _, n = orc.detection(rois, cls)
        t.append(time.perf_counter())
        pooled14, lv = orc.pyramid_roialign(det, maps, 14, size, size)
        t.append(time.perf_counter())
        nv = max(int((lv >= 0).sum()), 1)                # removeZeros: the Mask model only runs on valid blocks
        mk = ref.mask(np.ascontiguousarray(pooled14[:nv].transpose(0, 2, 3, 1))).numpy()
        full = np.zeros((100,) + mk.shape[1:], np.float32); full[:nv] = mk
        masks = orc.mask_select(full, (lv >= 0).astype(np.int32), det)
        t.append(time.perf_counter())
    names = ["backbone+fpn+rpn", "proposal", "roialign7", "classifier", "detection", "roialign14", "mask"]
    st["last_outputs"] = (img, det, masks)          # for the end-to-end agreement check (PipelineWorkload.cpu_baseline)
    return t[-1] - t[0], dict(zip(names, np.diff(t).tolist()))


def box_iou(a, b):
    """IoU of (y1, x1, y2, x2) boxes a [n, 4] against b [m, 4] (float64)."""
    a, b = a.astype(np.float64), b.astype(np.float64)
    iy = np.clip(np.minimum(a[:, None, 2], b[None, :, 2]) - np.maximum(a[:, None, 0], b[None, :, 0]), 0, None)
    ix = np.clip(np.minimum(a[:, None, 3], b[None, :, 3]) - np.maximum(a[:, None, 1], b[None, :, 1]), 0, None)
    inter = iy * ix
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]); ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / np.maximum(aa[:, None] + ab[None, :] - inter, 1e-30)


def detection_agreement(cpu_det, cpu_masks, gpu_det, gpu_masks, iou_thr=0.99):
    """The reference validates end to end (EvaluateCommand.swift:166-194, COCOEval/task.py:99-105): here, per image, every
    detection of the fp32 CPU path is matched one-to-one with a GPU detection of the same class and IoU > iou_thr."""
    nc, ng = int((cpu_det[:, 5] > 0).sum()), int((gpu_det[:, 5] > 0).sum())
    out = {"cpu": nc, "gpu": ng, "matched": 0, "max_box_delta": 0.0, "max_score_delta": 0.0, "max_mask_delta": 0.0}
    if nc == 0 or ng == 0:
        return out
    iou = box_iou(cpu_det[:nc, :4], gpu_det[:ng, :4])
    iou[cpu_det[:nc, 4][:, None] != gpu_det[:ng, 4][None, :]] = 0.0
    used = np.zeros(ng, bool)
    for i in np.argsort(-cpu_det[:nc, 5], kind="stable"):
        cand = np.where(~used & (iou[i] > iou_thr))[0]
        if len(cand) == 0:
            continue
        j = cand[np.argmax(iou[i, cand])]
        used[j] = True
        out["matched"] += 1
        out["max_box_delta"] = max(out["max_box_delta"], float(np.abs(cpu_det[i, :4] - gpu_det[j, :4]).max()))
        out["max_score_delta"] = max(out["max_score_delta"], float(abs(cpu_det[i, 5] - gpu_det[j, 5])))
        out["max_mask_delta"] = max(out["max_mask_delta"], float(np.abs(cpu_masks[i] - gpu_masks[j]).max()))
    return out


def run_reference(args):
    """--impl reference: the CPU path of the reference for this workload, timed on the host cores."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    import maskrcnn_b200 as m
    from oracle import oracle as orc
    orc.lib()
    cores = os.cpu_count() or 1
    workload = args.workload
    times = []
    arch, size, props, cfg_batch = CONFIGS[args.config]
    if workload == "pipeline":
        import torch
        torch.set_num_threads(cores)
        for s in range(args.warmup + args.steps):
            dt, _ = cpu_pipeline_image(m, orc, s, arch, size, props)
            if s >= args.warmup:
                times.append(dt)
        sample = (f"1 image ({size}x{size}, ResNet{arch}+FPN, {props} rois) per step: PyTorch-CPU fp32 dense graphs on all host "
                  "threads + oracle.c custom layers (1 thread)")
        kind_cores = cores
    else:
        workload = "custom_layers"
        anchors = m.synth.generate_anchors(IMG, IMG)
        maps = m.synth.feature_maps(0)
        for s in range(args.warmup + args.steps):
            dt = cpu_custom_layers_image(orc, m.synth, anchors, s, maps)
            if s >= args.warmup:
                times.append(dt)
        sample = "1 image per step: oracle custom layers (single thread, like the reference's layers)"
        kind_cores = 1
    ms = 1e3 * float(np.mean(times))
    v = 1e3 / ms
    emit(({
        "impl": "reference", "metric": config_metric(args.config), "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        # the same workload as the B200 arm's `config` (one image of that batch per step: a bounded sample)
        "config": {"workload": workload, "image": f"{size}x{size}x3", "batch_per_gpu": args.batch, "global_batch": args.batch * max(args.gpus, 1),
                   "name": args.config, "pre_nms": 6000, "rois": props, "detections": 100,
                   "model": f"ResNet{arch}+FPN Mask-RCNN, 81 classes, synthetic fp16 weights (seed 7)", "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": kind_cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# --------------------------------------------------------------------------- GPU workloads
class CustomLayersWorkload:
    """The five custom layers of the reference on synthetic RPN outputs / feature maps (no dense
    graph): ProposalLayer -> PyramidROIAlign(7) -> classifier select -> DetectionLayer ->
    PyramidROIAlign(14) -> Detection decode.  Layer-level ABI, reference (CHW fp32) layouts."""
    name = "custom_layers"
    dtype = "f32"

    def __init__(self, m, torch, device, batch, seed_base):
        self.m, self.torch, self.b = m, torch, batch
        self.ctx = m.Context(device=device, max_batch=batch)
        self.stream = torch.cuda.Stream()
        self.ctx.set_stream(self.stream.cuda_stream)
        anchors = m.synth.generate_anchors(IMG, IMG)
        self.ctx.set_anchors(anchors)
        # two distinct images' worth of RPN outputs, tiled (generation is slow on the host)
        gen = [m.synth.rpn_outputs(anchors, seed_base + i) for i in range(min(batch, 2))]
        probs = np.stack([gen[i % len(gen)][0] for i in range(batch)])
        deltas = np.stack([gen[i % len(gen)][1] for i in range(batch)])
        cls = [m.synth.classifier_outputs(1000, seed_base + i) for i in range(min(batch, 2))]
        cprob = np.stack([cls[i % len(cls)][0] for i in range(batch)])
        cbbox = np.stack([cls[i % len(cls)][1] for i in range(batch)])
        pin = lambda a: torch.from_numpy(a).pin_memory()
        self.h = {"probs": pin(probs), "deltas": pin(deltas), "cprob": pin(cprob), "cbbox": pin(cbbox)}
        g = torch.Generator(device="cuda").manual_seed(20260 + seed_base)
        self.d_maps = [torch.randn((batch, 256, IMG // s, IMG // s), device="cuda", generator=g) for s in (4, 8, 16, 32)]
        self.h_maps = None
        self.d = {k: v.cuda() for k, v in self.h.items()}
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device="cuda")
        self.out = {"rois": z(batch, 1000, 4), "pooled": z(batch, 1000, 256, 7, 7), "cls": z(batch, 1000, 6),
                    "det": z(batch, 100, 6), "pooled14": z(batch, 100, 256, 14, 14),
                    "cnt": z(batch, dt=torch.int32), "idx": z(batch, 100, dt=torch.int32),
                    "bbox": z(batch, 100, 4, dt=torch.float64), "dcls": z(batch, 100, dt=torch.int32),
                    "score": z(batch, 100, dt=torch.float64)}
        self.h_out = {k: torch.zeros_like(v, device="cpu").pin_memory() for k, v in self.out.items() if k in ("det", "cnt")}
        self.layers = (m.ProposalLayer(context=self.ctx), m.PyramidROIAlignLayer({"poolSize": 7}, context=self.ctx),
                       m.TimeDistributedClassifierLayer(context=self.ctx), m.DetectionLayer(context=self.ctx),
                       m.PyramidROIAlignLayer({"poolSize": 14}, context=self.ctx))
        self.h2d = sum(v.numel() * v.element_size() for v in self.h.values()) + sum(x.numel() * 4 for x in self.d_maps)
        self.d2h = sum(v.numel() * v.element_size() for v in self.h_out.values())

    def step(self, src=None, maps=None):
        src = src or self.d
        maps = maps or self.d_maps
        o = self.out
        prop, ra7, tdc, det, ra14 = self.layers
        prop.evaluate([src["probs"], src["deltas"]], [o["rois"]])
        ra7.evaluate([o["rois"]] + maps, [o["pooled"]])
        tdc.select(src["cprob"], src["cbbox"], o["cls"])
        det.evaluate([o["rois"], o["cls"]], [o["det"]])
        ra14.evaluate([o["det"]] + maps, [o["pooled14"]])
        self.m._cabi.check(self.ctx.handle, self.m.lib().mrcnn_detections_decode(
            self.ctx.handle, self.b, o["det"].data_ptr(), None, o["cnt"].data_ptr(), o["idx"].data_ptr(),
            o["bbox"].data_ptr(), o["dcls"].data_ptr(), o["score"].data_ptr(), None))

    def step_e2e(self):
        torch = self.torch
        if self.h_maps is None:
            self.h_maps = [x.cpu().pin_memory() for x in self.d_maps]
            self.e_maps = [torch.empty_like(x) for x in self.d_maps]
            self.e = {k: torch.empty_like(v) for k, v in self.d.items()}
        for k, v in self.h.items():
            self.e[k].copy_(v, non_blocking=True)
        for a, b in zip(self.e_maps, self.h_maps):
            a.copy_(b, non_blocking=True)
        self.step(self.e, self.e_maps)
        for k, v in self.h_out.items():
            v.copy_(self.out[k], non_blocking=True)
        self.stream.synchronize()

    def launches_per_step(self):
        c0 = self.ctx.launch_count
        self.step()
        return self.ctx.launch_count - c0

    def roofline(self, prof, peaks):
        ms, n, work = prof["roialign"]
        ach = work / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"kernel": "roialign_chw_kernel (pool 7 + pool 14 launches)", "bound": "hbm", "achieved": ach,
                "peak": peaks["hbm"], "peak_source": peaks["source"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                "traffic": None, "launches": n, "avg_launch_ms": ms / max(n, 1),
                "algorithmic_bytes_per_launch": work / max(n, 1)}


class PipelineWorkload:
    """BASELINE.json configs[1]: batch of 1024x1024x3 u8 images -> MaskRCNN.predict (ResNet101+FPN+RPN, ProposalLayer
    6000 -> 1000, PyramidROIAlign, classifier head, DetectionLayer, mask head) through mrcnn_predict; with N > 1 ranks
    every rank predicts its own batch and one NCCL all-gather collects all detections / masks (configs[3])."""
    name = "pipeline"
    dtype = "f16"     # tensor-core operands fp16 (as the reference stores its weights), fp32 accumulate; custom layers f32/f64

    def __init__(self, m, torch, device, batch, rank, world, architecture=101, size=IMG, proposals=1000, precise_masks=True):
        self.m, self.torch, self.b, self.world = m, torch, batch, world
        self.arch, self.size, self.props = architecture, size, proposals
        IMG = size                                                # (shadows the module constant: everything below is per config)
        cfg = m.MaskRCNNConfig()
        cfg.preciseMasks = bool(precise_masks)
        cfg.architecture = "resnet101" if architecture == 101 else "resnet50"
        cfg.imageShape, cfg.maxProposals = (size, size, 3), proposals
        cfg.maxBatch = batch
        _, blobs = m.weights.synthetic_blobs(architecture)
        self.model = m.MaskRCNN(cfg, device=device, blobs=blobs, anchors=m.synth.generate_anchors(IMG, IMG))
        self.ctx = self.model.ctx
        self.stream = torch.cuda.Stream()
        self.ctx.set_stream(self.stream.cuda_stream)
        rng = np.random.default_rng(20260 + rank)
        img = rng.integers(0, 256, (batch, IMG, IMG, 3), dtype=np.uint8)
        self.h_img = torch.from_numpy(img).pin_memory()
        self.d_img = self.h_img.cuda()
        tot = batch * world
        self.d_det = torch.zeros((tot, 100, 6), device="cuda")
        self.d_mask = torch.zeros((tot, 100, 28, 28), device="cuda")
        self.h_det = torch.zeros((tot, 100, 6)).pin_memory()
        self.h_mask = torch.zeros((tot, 100, 28, 28)).pin_memory()
        # streaming e2e: two batches in flight, each with its own pinned input / output buffers
        self.h_img2 = [self.h_img, self.h_img.clone().pin_memory()]
        self.h_det2 = [self.h_det, torch.zeros((tot, 100, 6)).pin_memory()]
        self.h_mask2 = [self.h_mask, torch.zeros((tot, 100, 28, 28)).pin_memory()]
        self.h2d = self.h_img.numel()
        self.d2h = 4 * (self.h_det.numel() + self.h_mask.numel())
        self.config_extra = {"model": f"ResNet{architecture}+FPN Mask-RCNN, 81 classes, synthetic fp16 weights (seed 7)",
                             "collective": "none" if world == 1 else "1 ncclAllGather of packed detections|masks per step",
                             "mask_head": "2-term fp16 activations: masks <= 1e-4 from an fp32 evaluation of the head (library default)" if precise_masks else
                                          "1-term fp16 activations (precise_masks = 0): masks within 5e-4 of fp32"}
        if world > 1:
            import torch.distributed as dist
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                buf = (C.c_char * 128)()
                m._cabi.check(None, m.lib().mrcnn_nccl_unique_id(buf))
                uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
            uid = uid.cuda()
            dist.broadcast(uid, 0)
            raw = bytes(uid.cpu().numpy().tobytes())
            m._cabi.check(self.ctx.handle, m.lib().mrcnn_comm_init(self.ctx.handle, raw, rank, world))

    def _predict(self, img, det, mask):
        l, h = self.m.lib(), self.ctx.handle
        if self.world > 1:
            self.m._cabi.check(h, l.mrcnn_predict_allgather(h, self.b, img.data_ptr(), det.data_ptr(), mask.data_ptr()))
        else:
            self.m._cabi.check(h, l.mrcnn_predict(h, self.b, img.data_ptr(), det.data_ptr(), mask.data_ptr()))

    def step(self):
        self._predict(self.d_img, self.d_det, self.d_mask)

    def verify_gather(self, dist, rank):
        """Outside the timed region, on hardware: slice r of the all-gathered detections / masks equals rank r's own
        mrcnn_predict of its images, bit for bit, and every rank holds identical gathered buffers (checksums all-reduced)."""
        torch = self.torch
        l, h, b = self.m.lib(), self.ctx.handle, self.b
        with torch.cuda.stream(self.stream):
            self._predict(self.d_img, self.d_det, self.d_mask)                       # gathered: [world * b, ...]
            ldet = torch.zeros((b, 100, 6), device="cuda"); lmask = torch.zeros((b, 100, 28, 28), device="cuda")
            self.m._cabi.check(h, l.mrcnn_predict(h, b, self.d_img.data_ptr(), ldet.data_ptr(), lmask.data_ptr()))
        self.stream.synchronize()
        mine = bool(torch.equal(self.d_det[rank * b:(rank + 1) * b], ldet) and torch.equal(self.d_mask[rank * b:(rank + 1) * b], lmask))
        nonzero = bool((ldet[..., 5] > 0).any())
        cs = (self.d_det.view(torch.int32).to(torch.int64).sum() * 31 + self.d_mask.view(torch.int32).to(torch.int64).sum()).reshape(1)
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok = torch.tensor([int(mine and nonzero and int(lo.item()) == int(hi.item()))], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        return bool(ok.item())

    def step_e2e(self):
        # the reference-facing call with HOST buffers: H2D of the images and D2H of detections + masks happen inside
        self._predict(self.h_img, self.h_det, self.h_mask)

    def run_e2e_stream(self, steps):
        """`steps` batches through mrcnn_predict_submit / mrcnn_predict_wait with HOST buffers: the H2D copy of batch
        i+1 overlaps the compute of batch i; every batch's H2D and D2H happen inside the loop, and the loop returns
        only after the last batch's results are in host memory."""
        l, h, chk = self.m.lib(), self.ctx.handle, self.m._cabi.check
        flags = 1 if self.world > 1 else 0
        for i in range(steps):
            k = i & 1
            chk(h, l.mrcnn_predict_submit(h, self.b, self.h_img2[k].data_ptr(), self.h_det2[k].data_ptr(),
                                          self.h_mask2[k].data_ptr(), flags))
            if i >= 1:
                chk(h, l.mrcnn_predict_wait(h))
        chk(h, l.mrcnn_predict_wait(h))

    def run_stream_device(self, steps):
        """`steps` batches through mrcnn_predict_submit / mrcnn_predict_wait with DEVICE-resident images and outputs (two
        batches in flight: the backbone of batch i + 1 runs beside the proposal / heads section of batch i)."""
        l, h, chk = self.m.lib(), self.ctx.handle, self.m._cabi.check
        flags = 1 if self.world > 1 else 0
        if not hasattr(self, "d_det2"):
            self.d_det2 = [self.d_det, self.torch.zeros_like(self.d_det)]
            self.d_mask2 = [self.d_mask, self.torch.zeros_like(self.d_mask)]
        for i in range(steps):
            k = i & 1
            chk(h, l.mrcnn_predict_submit(h, self.b, self.d_img.data_ptr(), self.d_det2[k].data_ptr(), self.d_mask2[k].data_ptr(), flags))
            if i >= 1:
                chk(h, l.mrcnn_predict_wait(h))
        chk(h, l.mrcnn_predict_wait(h))

    def launches_per_step(self):
        c0 = self.ctx.launch_count
        self.step()
        self.ctx.synchronize()
        return self.ctx.launch_count - c0

    def roofline(self, prof, peaks):
        ms, n, work = prof["conv_gemm_tcgen05"]
        ach = work / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        peak = peaks["bf16_sustained"]
        return {"kernel": "conv_gemm_kernel + conv_fused_expand_reduce_kernel (tcgen05 implicit-GEMM convolutions: every dense layer)", "bound": "tensor",
                "achieved": ach, "peak": peak, "peak_source": peaks["source"] + " (sustained cuBLAS bf16; fp16 runs at the same rate)",
                "unit": "TFLOP/s", "frac": ach / peak, "traffic": load_traffic("conv_gemm_tcgen05"), "traffic_unit": "B/launch (ncu dram bytes, profiles/traffic_r2.json)",
                "launches": n, "avg_launch_ms": ms / max(n, 1), "algorithmic_flops_per_launch": work / max(n, 1)}

    def extra_rooflines(self, prof, peaks):
        ms, n, work = prof["roialign"]
        ach = work / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        # `work` = output + roi bytes only: the map bytes a launch reads depend on the rois (the proposals of this step) and
        # are not known on the host; the kernel's roofline is measured by `bench.py --workload roialign` (configs[4])
        return {"roialign_in_pipeline": {"kernel": "roialign_staged_kernel (pool 7 + pool 14 launches)", "ms_per_step": ms / max(self._steps, 1),
                                         "launches": n, "output_and_roi_bytes_per_launch": work / max(n, 1),
                                         "output_GBps": ach, "note": "lower bound of the traffic (maps not counted); roofline: --workload roialign"},
                "stage_ms": dict(self.ctx.stage_times())}

    def cpu_baseline(self):
        from oracle import oracle as orc
        import torch
        orc.lib()
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        arch, size = self.arch, self.size
        cpu_pipeline_image(self.m, orc, 1000, arch, size, self.props)           # warm-up (weights to torch, page-in)
        runs, agree = [], []
        for i in range(4):
            runs.append(cpu_pipeline_image(self.m, orc, i, arch, size, self.props))
            # the SAME image through the GPU path: how far is the fp16-activation tensor-core pipeline from the fp32 CPU one?
            img, cdet, cmask = _CPU_STATE["last_outputs"]
            gdet, gmask = self.model.prediction_batch(img)
            agree.append(detection_agreement(cdet, cmask, gdet[0], gmask[0]))
        dt = float(np.mean([r[0] for r in runs]))
        stages = {k: float(np.mean([r[1][k] for r in runs])) for k in runs[0][1]}
        tot = {k: (sum(a[k] for a in agree) if k in ("cpu", "gpu", "matched") else max(a[k] for a in agree)) for k in agree[0]}
        self.e2e_agreement = dict(tot, images=len(agree), rule="same class and IoU > 0.99, one to one; deltas over the matched detections",
                                  note="GPU: fp16 activations in backbone / classifier head (2-term in the mask head); CPU: fp32 everywhere, same fp16-stored weights")
        return {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "4 images of the same workload after 1 warm-up image: PyTorch-CPU fp32 dense graphs (all host threads) "
                          "+ oracle.c custom layers (1 thread)", "stage_seconds": stages}


# --------------------------------------------------------------------------- ROIAlign microbench (configs[4])
ROI_METRIC = "ROIAlign HBM GB/s (1000 ROIs x P2 256x256x256 tiles)"
ROI_HEADLINE = ("nhwc_f16", 8, 1000, 7)


def roialign_cpu(orc, batch_images=1, r=1000, pool=7, reps=3):
    """The reference's ROIAlign on the host CPU: oracle.c crop_and_resize (scalar, one thread, CHW fp32 = the layer's own
    layout) on the same level-2 rois; returns (seconds per image, bytes per image by the touched model)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_roialign as br
    rng = np.random.default_rng(5)
    maps = [rng.standard_normal((256, s, s), dtype=np.float32) for s in (256, 128, 64, 32)]
    ts, byts = [], 0
    for i in range(batch_images * reps):
        rois = br.level2_rois(r, 100 * i + r)
        t0 = time.perf_counter()
        orc.pyramid_roialign(rois, maps, pool)
        ts.append(time.perf_counter() - t0)
        byts = br.touched_map_bytes(rois, 256, 256, 256, pool, "chw_f32") + r * 256 * pool * pool * 4 + r * 16
    return float(np.median(ts)), byts


def run_roialign(args):
    """BASELINE.json configs[4].  One "step" = one ROIAlign launch over the batch (level kernel + ROIAlign kernel).  value
    = GB/s of the headline case (internal NHWC fp16 layout, batch 8, 1000 rois, pool 7) on the TOUCHED-bytes model
    (tools/bench_roialign.py); the SURVEY 8(d) whole-map model and the ncu DRAM bytes are reported beside it."""
    rank, world, local = dist_env()
    if rank != 0:
        return                                                # independent replicas would measure the same thing
    import torch
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    import maskrcnn_b200 as m
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_roialign as br
    peaks = load_peaks()
    sampler = ClockSampler(local)
    sampler.start()
    cases = None
    if args.quick:
        cases = [ROI_HEADLINE, ("nhwc_f16", 8, 1000, 14), ("nhwc_f16", 1, 1000, 7), ("chw_f32", 1, 1000, 7), ("chw_f32", 8, 1000, 7)]
    rows = br.sweep(m, torch, cases, iters=max(args.steps, 4), log=lambda s: sys.stderr.write(s + "\n"))
    clocks = sampler.stop()
    head = next(r for r in rows if (r["layout"], r["batch"], r["rois"], r["pool"]) == ROI_HEADLINE)
    chw = next((r for r in rows if (r["layout"], r["batch"], r["rois"], r["pool"]) == ("chw_f32", 8, 1000, 7)), None)
    # end to end: the layer call with HOST buffers (maps, rois in; pooled features out), copies inside the timed region
    ctx = m.Context()
    st = torch.cuda.Stream(); ctx.set_stream(st.cuda_stream)
    b, r, pool = 8, 1000, 7
    hmaps = [torch.randn((b, 256, s, s)).pin_memory() for s in (256, 128, 64, 32)]
    hrois = torch.from_numpy(np.stack([br.level2_rois(r, 100 * i + r) for i in range(b)])).pin_memory()
    hout = torch.empty((b, r, 256, pool, pool)).pin_memory()
    layer = m.PyramidROIAlignLayer({"poolSize": pool}, context=ctx)
    ts = []
    for it in range(4):
        t0 = time.perf_counter()
        layer.evaluate([hrois] + hmaps, [hout])
        ts.append(time.perf_counter() - t0)
    e2e_s = float(np.median(ts[1:]))
    h2d = sum(x.numel() * 4 for x in hmaps) + hrois.numel() * 4
    d2h = hout.numel() * 4
    e2e_bytes = sum(br.touched_map_bytes(hrois[i].numpy(), 256, 256, 256, pool, "chw_f32") for i in range(b)) + d2h + hrois.numel() * 4
    ctx.close()
    line = {
        "metric": ROI_METRIC, "value": head["GBps_touched"], "unit": "GB/s", "n_gpus": 1, "steps": max(args.steps, 4),
        "warmup": 2, "ms_per_step": head["us_call"] * 1e-3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 maps / f32 arithmetic (internal NHWC layout); f32 (boundary CHW layout)", "data": "synthetic",
        "config": {"workload": "roialign", "case": head["case"], "fmap": "P2 256x256x256 per image", "rois_per_image": 1000,
                   "batch": 8, "pool": 7, "rois": "level 2 only: sqrt(w*h) in [12, 75] px, log-uniform, seeded",
                   "l2_policy": "256 MB memset before every timed launch"},
        "clocks": clocks,
        "e2e": {"value": e2e_bytes / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "mode": "mrcnn_pyramid_roialign_eval (boundary CHW fp32 layout) with pinned HOST maps / rois / output, one blocking "
                        "call per step; the 716 MB of feature maps cross PCIe every call, as they cross to the GPU in the reference "
                        "(PyramidROIAlignLayer.swift:110-118)", "seconds_per_step": e2e_s},
        "gpu_launches": 2 * (max(args.steps, 4) + 2),
        "roofline": {"kernel": "roialign_staged_kernel<7, 7, 2, false, false> (NHWC fp16)", "bound": "hbm", "achieved": head["GBps_touched"],
                     "peak": peaks["hbm"], "peak_source": peaks["source"], "unit": "GB/s", "frac": head["frac"],
                     "algorithmic_bytes_per_launch": head["bytes_touched"], "avg_launch_us": head["us_kernel"],
                     "achieved_survey_model": head["GBps_survey"], "frac_survey_model": head["frac_survey"],
                     "achieved_dram": head["GBps_dram_ncu"], "frac_dram": head["frac_dram_ncu"], "traffic": head["bytes_dram_ncu"],
                     "note": "achieved = (map bytes the rois touch, 32-B sectors + output + rois) / kernel time; survey model charges "
                             "the whole P2 map (SURVEY.md 8(d)); achieved_dram = ncu dram__bytes of the same launch / kernel time"},
        "roofline_chw_f32": None if chw is None else {"kernel": "roialign_staged_kernel<7, 7, 2, false, true> (boundary CHW fp32)", "bound": "hbm", "achieved": chw["GBps_touched"],
                                                      "peak": peaks["hbm"], "unit": "GB/s", "frac": chw["frac"], "frac_survey_model": chw["frac_survey"],
                                                      "achieved_dram": chw["GBps_dram_ncu"], "frac_dram": chw["frac_dram_ncu"], "avg_launch_us": chw["us_kernel"]},
        "sweep": rows,
    }
    if not args.no_cpu_baseline:
        from oracle import oracle as orc
        orc.lib()
        sec, byts = roialign_cpu(orc)
        line["cpu_baseline"] = {"value": byts / sec / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                                "sample": "1 image (1000 level-2 rois, pool 7, CHW fp32) x 3, oracle.c crop_and_resize, single thread "
                                          "like the reference's layer code", "seconds_per_image": sec}
    emit(line)


def run_roialign_reference(args):
    """The reference's ROIAlign on the host: oracle.c crop_and_resize, one image per host thread (the layer itself is
    single-threaded; images are independent, so all cores are used by running `cores` images side by side)."""
    rank, _, _ = dist_env()
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_roialign as br
    orc.lib()
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(5)
    maps = [rng.standard_normal((256, s, s), dtype=np.float32) for s in (256, 128, 64, 32)]
    rois = [br.level2_rois(1000, 100 * i + 1000) for i in range(cores)]
    byts = sum(br.touched_map_bytes(r, 256, 256, 256, 7, "chw_f32") + 1000 * 256 * 49 * 4 + 16000 for r in rois)
    ts = []
    with ThreadPoolExecutor(cores) as ex:
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            list(ex.map(lambda r: orc.pyramid_roialign(r, maps, 7), rois))       # ctypes releases the GIL
            if s >= args.warmup:
                ts.append(time.perf_counter() - t0)
    sec = float(np.mean(ts))
    v = byts / sec / 1e9
    sample = (f"{cores} images per step, one per host thread (1000 level-2 rois each, pool 7, CHW fp32): oracle.c crop_and_resize")
    emit({"impl": "reference", "metric": ROI_METRIC, "value": v, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f32", "data": "synthetic",
          "config": {"workload": "roialign", "case": "chw_f32,1,1000,7", "fmap": "P2 256x256x256 per image", "rois_per_image": 1000,
                     "batch": 8, "pool": 7, "sample": sample},
          "cpu_baseline": {"value": v, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
          "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def make_workload(args, m, torch, device):
    rank, world, _ = dist_env()
    if args.workload == "pipeline":
        arch, size, props, _ = CONFIGS[args.config]
        return PipelineWorkload(m, torch, device, args.batch, rank, world, architecture=arch, size=size, proposals=props,
                                precise_masks=args.precise_masks)
    return CustomLayersWorkload(m, torch, device, args.batch, 0)


def run_ours(args):
    import torch
    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    import maskrcnn_b200 as m
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = make_workload(args, m, torch, local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, whole=False):
        # CUDA events on the stream the library launches on (wl.stream is the context's stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(wl.stream):
            e0.record()
            if whole:
                fn(steps)           # runs all `steps` itself (streaming submit / wait loop)
            else:
                for _ in range(steps):
                    fn()
            e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    with torch.cuda.stream(wl.stream):
        launches = wl.launches_per_step()
        for _ in range(max(args.warmup, 3)):
            wl.step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    wl.ctx.profile_enable(False)
    total_ms = timed(wl.step, args.steps)                    # one blocking mrcnn_predict per step, device-resident inputs
    value_mode = "one blocking call per step (mrcnn_predict), device-resident images and outputs"
    stream_dev_ms = None
    if hasattr(wl, "run_stream_device"):
        with torch.cuda.stream(wl.stream):
            wl.run_stream_device(3)
        stream_dev_ms = timed(wl.run_stream_device, args.steps, whole=True)
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel-class timing pass (event pairs perturb the stream slightly -> separate from `value`)
    wl.ctx.profile_enable(True)
    barrier()
    with torch.cuda.stream(wl.stream):
        for _ in range(args.steps):
            wl.step()
    prof = wl.ctx.profile_read()
    wl.ctx.profile_enable(False)
    wl._steps = args.steps
    # end to end through host buffers
    with torch.cuda.stream(wl.stream):
        for _ in range(2):
            wl.step_e2e()
    e2e_sync_ms = timed(wl.step_e2e, args.steps)
    e2e_ms, e2e_mode = e2e_sync_ms, "one synchronous call per step"
    if hasattr(wl, "run_e2e_stream"):
        with torch.cuda.stream(wl.stream):
            wl.run_e2e_stream(2)
        e2e_ms = timed(wl.run_e2e_stream, args.steps, whole=True)
        e2e_mode = ("streaming: mrcnn_predict_submit / mrcnn_predict_wait, two batches in flight, pinned host buffers; "
                    "every step's H2D and D2H inside the timed region, which ends after the last batch's results are in host memory")

    gather_verified = None
    if world > 1:
        if hasattr(wl, "verify_gather"):
            gather_verified = wl.verify_gather(dist, rank)
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    ms_step = total_ms / args.steps
    imgs = args.batch * world
    line = {
        "metric": config_metric(args.config), "value": imgs / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "value_mode": value_mode,
        # the same with the streaming calls (two batches in flight, heads of batch i beside the backbone of batch i + 1): the
        # board is power-capped during the step (clocks.reasons), so overlap buys little at N = 1; at N > 1 it hides the all-gather
        "streaming_device_value": imgs / (stream_dev_ms / args.steps * 1e-3) if stream_dev_ms else None,
        "config": dict({"workload": wl.name, "name": args.config, "image": f"{CONFIGS[args.config][1]}x{CONFIGS[args.config][1]}x3",
                        "batch_per_gpu": args.batch, "global_batch": imgs,
                        "pre_nms": 6000, "rois": CONFIGS[args.config][2], "detections": 100,
                        "l2_policy": "per-step working set (batch activations + feature maps, > 1 GB) exceeds the 126 MB L2; no explicit flush"},
                       **getattr(wl, "config_extra", {})),
        "clocks": clocks,
        "e2e": {"value": imgs / (e2e_ms / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": wl.h2d * world,
                "d2h_bytes_per_step": wl.d2h * world, "mode": e2e_mode,
                "sync_value": imgs / (e2e_sync_ms / args.steps * 1e-3),
                "sync_mode": "mrcnn_predict with host buffers, one blocking call per step (copies serialised with compute)"},
        "gpu_launches": launches * args.steps,
        "gather_verified": gather_verified,       # N > 1: gathered slice r == rank r's local predict, identical buffers on all ranks
        "roofline": wl.roofline(prof, peaks),
        "kernel_classes": {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps,
                               "work_per_step": v[2] / args.steps} for k, v in prof.items() if v[1]},
    }
    if hasattr(wl, "extra_rooflines"):
        line.update(wl.extra_rooflines(prof, peaks))
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = wl.cpu_baseline() if hasattr(wl, "cpu_baseline") else default_cpu_baseline(m)
        if hasattr(wl, "e2e_agreement"):
            line["e2e_agreement"] = wl.e2e_agreement
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def default_cpu_baseline(m):
    from oracle import oracle as orc
    orc.lib()
    anchors = m.synth.generate_anchors(IMG, IMG)
    maps = m.synth.feature_maps(0)
    ts = [cpu_custom_layers_image(orc, m.synth, anchors, i, maps) for i in range(3)]
    dt = float(np.mean(ts[1:]))
    return {"value": 1.0 / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "2 images (after 1 warm-up) of the same custom-layer workload, oracle/liboracle.so, single thread"}


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter) was re-routed to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # C-level stdout of this process (e.g. "NCCL version ...") -> stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "pipeline", "custom_layers", "roialign"])
    ap.add_argument("--quick", action="store_true", help="roialign: the headline cases only instead of the whole sweep")
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default 8 = configs[1]; 1 for --config single)")
    ap.add_argument("--config", default="A", choices=sorted(CONFIGS), help="A: ResNet101 1024x1024 1000 rois (headline); small: ResNet50 "
                    "512x512 300 rois (configs[2]); single: A with one image per call (configs[0], latency)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precise-masks", type=int, default=1, help="1 (default): 2-term fp16 activations in the mask head, masks within 1e-4 of fp32; 0: 1-term, 5e-4")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "pipeline"
    if args.batch is None:
        args.batch = CONFIGS[args.config][3] or 8
    if args.workload == "roialign":
        (run_roialign_reference if args.impl == "reference" else run_roialign)(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
