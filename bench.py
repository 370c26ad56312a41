#!/usr/bin/env python
"""bench.py -- stereo UVO frames/sec on B200 (BASELINE.json config[1]: stereo 1280x1024, ~4k SURF features/image,
3D-to-2D solvePnPRansac), with the kernel roofline and the CPU baseline beside it.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (the CPU path -- the oracle port -- on the host cores)
  python bench.py --config E ...                           (BASELINE config[4]: 2448x2048, one sequence per GPU)

A step = one stereo frame through the whole hot path (get_image x2 -> SURF x2 -> stereo match -> temporal match ->
triangulate -> extract_3Dpoints -> solvePnPRansac -> velocity).
  value : frames/s with the input images already resident in HBM (device-resident ring larger than L2), frames
          enqueued asynchronously through the C ABI, timed with CUDA events on the library's stream.
  e2e   : frames/s through the reference-facing host-buffer call (uvo_stereo_enqueue_host + uvo_stereo_collect) with
          HOST (pinned) images: H2D of both images and D2H of the result record inside the timed region.
Both are the MEDIAN of --regions (9) consecutive timed regions of exactly K frames, each bracketed by barrier +
synchronize, max over ranks per region; `timing` carries every region and one long steady-state region.
  roofline : the top kernel by time per frame over ALL kernels (each timed alone with CUDA events); HBM, tensor or --
          for the fp64 pose kernels -- the FP64-issue bound measured live by tools/probes/fp64_probe.
Multi-GPU: VO is sequential within a stream, so ranks run independent sequences (seed 1300+rank) -- replicas,
no collective on the data path; torch.distributed is used only for the barrier and the max-over-ranks of the time.
"""
import argparse
import json
import os

# more hardware work queues than the default 8: the stereo handle uses 8 lane streams + a copy stream + the ingest
# streams of compressed input; set before CUDA initialises (torch does that first in this process)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TARGET_KP = 4096
# BASELINE.json configs: B = configs[1] (the configuration the metric is quoted on, the default and the headline),
# E = configs[4] (2448x2048, one sequence per GPU; run it at N = 1, 2, 4, 8 with --config E).
#   n_distinct: rendered stereo pairs, the sequence ping-pongs through them (0..n-1, n-2..1, 0..)
#   ring: device-resident pairs (whole ping-pong cycles, larger than the 126 MB L2)
CONFIGS = {
    "B": dict(w=1280, h=1024, n_distinct=8, ring=28,
              workload="stereo UVO 1280x1024, ~4k SURF features/image, solvePnPRansac(EPNP), shipped stereo YAML "
                       "(BASELINE config[1]); one independent sequence per GPU"),
    "E": dict(w=2448, h=2048, n_distinct=4, ring=16,
              workload="stereo UVO 2448x2048, ~4k SURF features/image, solvePnPRansac(EPNP), shipped stereo YAML "
                       "(BASELINE config[4]); one independent sequence per GPU"),
}
REGIONS = 9          # consecutive timed regions of `steps` frames each; the median region is the reported one


def pingpong(i, n):
    period = 2 * n - 2
    j = i % period
    return j if j < n else period - j


def clocks_sampler(stop, out, gpu_index):
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                              str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
        return
    def reader():
        for line in p.stdout:
            out.append(line.strip())
    t = threading.Thread(target=reader, daemon=True)
    t.start()
    stop.wait()
    p.terminate()
    t.join(timeout=2)


def summarise_clocks(lines):
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        f = [x.strip() for x in ln.split(",")]
        if len(f) < 7:
            continue
        try:
            sm.append(float(f[0]))
            mx.append(float(f[1]))
        except ValueError:
            continue
        for name, v in zip(names, f[3:7]):
            if v.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
            "samples": len(sm)}


def aggregate_over_ranks(dist, ms_list, count_list, device):
    """multi-GPU reduction used by the bench: MAX over ranks of each time, SUM over ranks of each count (replicas:
    one independent sequence per rank, no data-path collective).  Works with nccl (cuda) and gloo (cpu)."""
    import torch
    t = torch.tensor(ms_list, dtype=torch.float64, device=device)
    v = torch.tensor(count_list, dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    return t.tolist(), v.tolist()


def bind_to_gpu_numa_node(torch, local):
    """multi-GPU runs: keep this rank's threads -- and through first touch its pinned image ring -- on the NUMA node its
    GPU hangs off (sysfs), so that N ranks do not pull their H2D traffic across the socket link.  Best effort: returns
    a description, or None when the topology cannot be read."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        return None


def sequence_seed(rank):
    """independent camera sequence per rank (seeds 1300+g)"""
    return 1300 + int(rank)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def make_sequence(seed, cfg, n_frames=None):
    from tools import synth
    return synth.StereoSequence(cfg["w"], cfg["h"], n_frames=n_frames or cfg["n_distinct"], seed=seed, tex_size=2048)


def fp64_peak():
    """FP64-issue roofline of the pose kernels: sustained DFMA rate of this GPU, measured live by
    tools/probes/fp64_probe (built by __graft_entry__.build()); the committed round-2 measurement otherwise."""
    exe = os.path.join(ROOT, "tools", "probes", "fp64_probe")
    try:
        d = json.loads(subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout)
        return float(d["dfma_tflops"]), "measured live (tools/probes/fp64_probe: dependent DFMA %.1f cycles)" % \
            d["latency_cycles"]["dfma"]
    except Exception:
        pass
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "fp64_probe_r02.json")))
        return float(d["dfma_tflops"]), "profiles/fp64_probe_r02.json"
    except Exception:
        return 37.0, "fallback (B200 FP64 nominal)"


def pick_threshold(ctx, U, seq):
    """bisect the integer SURF threshold on frame 0 (GPU path) for TARGET_KP +- 2 % keypoints, then freeze it"""
    g = ctx.get_image(seq.frames[0][0], seq.KL, seq.DL, seq.newKL)
    lo, hi = 100, 200000
    ctx.params.max_features = 1 << 16
    best = None
    while lo < hi:
        mid = (lo + hi) // 2
        ctx.params.surf_min_hessian = mid
        n = len(ctx.detect_features(g)[0])
        if best is None or abs(n - TARGET_KP) < abs(best[1] - TARGET_KP):
            best = (mid, n)
        if abs(n - TARGET_KP) <= 0.02 * TARGET_KP:
            break
        if n > TARGET_KP:
            lo = mid + 1
        else:
            hi = mid
    return best


def pick_threshold_cpu(seq):
    """same rule as pick_threshold, evaluated with the CPU path (reference arm only)"""
    from oracle import oracle as O
    g = O.get_image(seq.frames[0][0], seq.KL, seq.DL, seq.newKL, True, 8.0)
    lo, hi, best = 100, 200000, None
    while lo < hi:
        mid = (lo + hi) // 2
        n = len(O.surf_detect_and_compute(g, mid)[0])
        if best is None or abs(n - TARGET_KP) < abs(best[1] - TARGET_KP):
            best = (mid, n)
        if abs(n - TARGET_KP) <= 0.02 * TARGET_KP:
            break
        if n > TARGET_KP:
            lo = mid + 1
        else:
            hi = mid
    return best[0]


# ---------------------------------------------------------------------------------------------------- CPU arm
def cpu_frames(seq, thr, n_frames, params=None):
    """the reference's CPU path (oracle port of the OpenCV calls) on n_frames consecutive frames; returns seconds per
    frame (steady state: the initialisation frame is not counted)"""
    from oracle import oracle as O
    from oracle.ref_stereo import RefStereoVO
    import ergo_uvo_b200._lib as L
    p = params or default_params_cpu(thr)
    ref = RefStereoVO(O, seq, p)
    ref.frame(*seq.frames[0], 0.1)  # initialisation frame
    ref.stage_s = {k: 0.0 for k in ref.stage_s}
    t0 = time.perf_counter()
    valid = 0
    for k in range(1, n_frames + 1):
        r = ref.frame(*seq.frames[pingpong(k, len(seq.frames))], 0.1)
        valid += r["valid"]
    dt = time.perf_counter() - t0
    cpu_frames.last_stage_ms = {k: 1e3 * v / n_frames for k, v in ref.stage_s.items()}
    cpu_frames.last_stage_ms_cv2 = cv2_stage_ms(seq, r)
    return dt / n_frames, valid


def cv2_stage_ms(seq, last):
    """the stages OpenCV's Python wheel can run (no contrib SURF), on the same frame: get_image for the pair and one
    BFMatcher knnMatch(k=2) on the frame's own descriptors -- the 'per-stage ms vs CPU OpenCV' column"""
    try:
        import cv2
    except ImportError:
        return None
    import numpy as np
    L, R = seq.frames[1 % len(seq.frames)]
    clahe = cv2.createCLAHE(clipLimit=8.0, tileGridSize=(8, 8))

    def get_image(img, K, D, newK):
        g = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)
        g = cv2.undistort(g, K, np.asarray(D, dtype=np.float64), None, newK)
        return clahe.apply(g)

    def best(fn, n=3):
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return 1e3 * min(ts)
    out = {"cv2": cv2.__version__, "threads": cv2.getNumThreads(),
           "get_image": best(lambda: (get_image(L, seq.KL, seq.DL, seq.newKL), get_image(R, seq.KR, seq.DR, seq.newKR)))}
    if last is not None and "dL" in last and len(last["dL"]) and len(last["dR"]):
        bf = cv2.BFMatcher(cv2.NORM_L2, False)
        dL, dR = np.ascontiguousarray(last["dL"]), np.ascontiguousarray(last["dR"])
        out["match_stereo"] = best(lambda: bf.knnMatch(dL, dR, 2), 2)
    return out


def default_params_cpu(thr):
    """uvo_params with the shipped stereo YAML values, without touching the CUDA library"""
    import ctypes as C
    from ergo_uvo_b200 import _lib as L
    p = L.Params()
    p.clahe, p.clip_limit, p.distance, p.lowe_ratio = 1, 8, 10, 0.8
    p.reprojection_tolerance, p.min_num_features, p.min_num_3dpoints, p.min_num_inliers = 3.0, 5, 5, 5
    p.iterations_count, p.reprojection_error, p.confidence, p.pnp_method_flag = 1000, 1.0, 0.99, 1
    p.surf_min_hessian, p.surf_octaves, p.surf_octave_layers, p.surf_extended, p.surf_upright = thr, 4, 3, 0, 1
    p.max_features = 16384
    return p


def config_dict(cfg, name, thr):
    """the `config` object of both arms (identical for the same --config and threshold, so that the driver can tell the
    two lines describe the same workload); what is specific to a run is in `run`"""
    P = cfg["w"] * cfg["h"]
    return {"workload": cfg["workload"], "name": name, "width": cfg["w"], "height": cfg["h"], "surf_min_hessian": thr,
            "l2": f"GPU arm: inputs larger than L2 -- a ring of {cfg['ring']} device-resident stereo pairs "
                  f"({cfg['ring'] * 2 * 3 * P / 1e6:.0f} MB), each read once per {cfg['ring']} frames; "
                  "CPU arm: consecutive frames of the same sequence from host memory"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = CONFIGS[args.config]
    seq = make_sequence(1300, cfg, n_frames=4)
    thr = args.threshold or pick_threshold_cpu(seq)
    cores = min(16, os.cpu_count() or 1)
    for _ in range(max(args.warmup, 0)):
        cpu_frames(seq, thr, 1)
    spf, valid = cpu_frames(seq, thr, max(args.steps, 1))
    fps = 1.0 / spf
    line = {
        "impl": "reference", "metric": "stereo_uvo_frames_per_sec", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": spf * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32/f32/f64",
        "data": "synthetic",
        "config": config_dict(cfg, args.config, thr),
        # one CPU process runs ONE sequence whatever --gpus says (the contract: rank 0 alone runs the reference arm);
        # a ratio against an N-GPU line compares N sequences with this one
        "n_sequences": 1,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} consecutive stereo frames of the same synthetic sequence through the "
                                   "oracle port of the OpenCV CPU path (SURF restated, not OpenCV)",
                         "stage_ms": getattr(cpu_frames, "last_stage_ms", None),
                         "stage_ms_cv2": getattr(cpu_frames, "last_stage_ms_cv2", None)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "valid_frames": valid,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------- GPU arm
def kernel_models(P, n_kp, n3d, iters):
    """ALGORITHMIC bytes / flops per launch of each kernel: name -> (bound, amount).
    hbm: bytes; tensor: flops of the TF32 contraction; fp64: double-precision flops (FP64-issue roofline)."""
    # EPnP on a minimal set, counted from the code: 12x12 round-robin Jacobi ~60 rounds x (78 x 13 + 144 x 3 + 6 x 60)
    # ~ 1.1e5, three beta systems + Gauss-Newton ~ 2.5e4, small SVDs / Rodrigues ~ 1e4
    f_epnp = 1.45e5
    f_score = 32.0  # PnP reprojection, flops per (hypothesis, correspondence)
    m = {
        # one launch covers both images of the pair (blockIdx.z = image)
        "k_gray_undistort": ("hbm", 2 * 4.0 * P), "k_clahe_tile_lut": ("hbm", 2 * 1.0 * P),
        "k_clahe_apply": ("hbm", 2 * 2.0 * P), "k_integral_rows": ("hbm", 2 * 5.0 * P),
        "k_integral_cols": ("hbm", 2 * 8.0 * P),
        "k_surf_detect": ("hbm", 2 * 16.0 * P), "k_surf_patch": ("hbm", 2 * n_kp * (40 * 40 + 256)),
        "k_surf_vector": ("hbm", 2 * n_kp * (441 + 256)),
        "k_knn_tc": ("tensor", 2.0 * n_kp * n_kp * 64),
        "k_triangulate": ("fp64", n3d * 2000.0),                      # 4x4 SVD ~ 2 kflop per point
        "k_pnp_chunk[0:32]": ("fp64", 32 * (n3d * f_score + f_epnp)),
        # refit: inlier test over all points + one EPnP on the inliers (40 sums + 3 x (9 + 27 + 3) sums per point)
        "k_pnp_finalize": ("fp64", n3d * (f_score + 2 * 40 + 3 * 2 * 39 + 60) + f_epnp),
    }
    return m


def run_gpu(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import ergo_uvo_b200 as U

    cfg = CONFIGS[args.config]
    W, H, N_DISTINCT, RING = cfg["w"], cfg["h"], cfg["n_distinct"], cfg["ring"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    affinity = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    if world > 1:
        # NCCL is used for the barrier / MAX-over-ranks only; its debug output (the "NCCL version ..." banner at
        # NCCL_DEBUG=VERSION and above) goes to stdout by default: send it to stderr so that stdout carries the one
        # JSON line only
        # (NCCL honours NCCL_DEBUG_FILE only above the VERSION level)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream(device=local)
    ctx = U.Context(local, stream=stream.cuda_stream)

    seq = make_sequence(sequence_seed(rank), cfg)
    if args.threshold:
        thr, n0 = args.threshold, -1
    else:
        thr, n0 = pick_threshold(ctx, U, seq)
    p = U.default_params(True)
    p.surf_min_hessian = thr
    p.max_features = 16384
    camL = U.make_camera(seq.KL, seq.DL, seq.newKL)
    camR = U.make_camera(seq.KR, seq.DR, seq.newKR)
    vo = U.StereoVO(ctx, W, H, camL, camR, seq.R_right, seq.t_right, p)
    vo.set_graphs(args.graphs)

    # device ring (inputs resident in HBM, larger than L2) and pinned host ring
    pitch = 3 * W
    dev_ring = [(torch.from_numpy(seq.frames[pingpong(i, N_DISTINCT)][0]).cuda(),
                 torch.from_numpy(seq.frames[pingpong(i, N_DISTINCT)][1]).cuda()) for i in range(RING)]
    host_ring = [(torch.from_numpy(seq.frames[pingpong(i, N_DISTINCT)][0]).pin_memory(),
                  torch.from_numpy(seq.frames[pingpong(i, N_DISTINCT)][1]).pin_memory()) for i in range(RING)]
    torch.cuda.synchronize()
    dt_frame = 0.1
    inflight = args.inflight if args.inflight > 0 else vo.lanes()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_enqueue_s = [0.0, 0]  # host time inside the enqueue call (launches, copies, event waits), calls; timed regions

    def run_frames(n, start, inflight=inflight, host=False, count_host_time=False):
        ring = host_ring if host else dev_ring
        enq = vo.enqueue_host if host else vo.enqueue_device
        valid = 0
        q = 0
        for i in range(n):
            L, R = ring[(start + i) % RING]
            t_enq = time.perf_counter()
            enq(L.data_ptr(), R.data_ptr(), pitch, dt_frame)
            if count_host_time:
                host_enqueue_s[0] += time.perf_counter() - t_enq
                host_enqueue_s[1] += 1
            q += 1
            if q >= inflight:
                valid += vo.collect().valid
                q -= 1
        while q:
            valid += vo.collect().valid
            q -= 1
        return valid

    def run_host_sync(n, start):
        valid = 0
        for i in range(n):
            L, R = host_ring[(start + i) % RING]
            r = U.StereoResult()
            rc = vo.lib.uvo_stereo_frame(vo.h, C.c_void_p(L.data_ptr()), C.c_void_p(R.data_ptr()), C.c_size_t(pitch),
                                         C.c_double(dt_frame), C.byref(r))
            ctx._ck(rc)
            valid += r.valid
        return valid

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pos = [0]

    def timed_region(n, host, count_host_time=False):
        """EXACTLY n frames between barrier + synchronize on both sides; CUDA events on the library's stream (wall clock
        as a floor for the host-buffer path, whose copies run on the library's own copy stream)"""
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record(stream)
            valid = run_frames(n, pos[0], host=host, count_host_time=count_host_time)
            e1.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        pos[0] += n
        return (max(ms, wall) if host else ms), valid

    # ---- device-resident throughput (value): REGIONS consecutive regions of `steps` frames, the median is reported
    n_warm = max(args.warmup, 3)
    n_roll = max(n_warm, 2 * vo.lanes() + 4)  # every lane has run a frame directly and captured its graphs
    run_frames(n_roll, pos[0])
    pos[0] += n_roll
    stop, samples = threading.Event(), []
    sampler = threading.Thread(target=clocks_sampler, args=(stop, samples, local), daemon=True)
    sampler.start()
    launches0, graphs0 = ctx.launches, vo.graph_launches
    dev_regions, valid_dev = [], 0
    for _ in range(args.regions):
        ms, v_ = timed_region(args.steps, host=False, count_host_time=True)
        dev_regions.append(ms)
        valid_dev += v_
    launches = ctx.launches - launches0
    graph_launches = vo.graph_launches - graphs0

    # ---- end-to-end through the host-buffer call (e2e)
    n_warm_host = max(args.warmup, vo.max_in_flight())  # every staging slot of the library's ring allocated and touched
    run_frames(n_warm_host, pos[0], host=True)
    pos[0] += n_warm_host
    host_regions, valid_host = [], 0
    for _ in range(args.regions):
        ms, v_ = timed_region(args.steps, host=True)
        host_regions.append(ms)
        valid_host += v_
    # ---- end-to-end with COMPRESSED input (the node's real input, sensor_msgs::CompressedImage): JPEG bytes in host
    # memory -> Huffman decode on `jpeg_threads` host threads (uvo_jpeg_entropy_decode_sparse, GIL released) -> sparse
    # coefficients over PCIe -> IDCT + colour on the frame's lane -> the same frame.  The encoding (the camera's
    # job) is done before the timed region.
    comp = None
    if args.jpeg_threads > 0:
        try:
            import cv2
            from concurrent.futures import ThreadPoolExecutor
            enc = []
            for (Li, Ri) in seq.frames:
                pair = []
                for img in (Li, Ri):
                    ok, e = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                                       cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420])
                    pair.append(np.ascontiguousarray(e).ravel())
                enc.append(tuple(pair))
            K_ = 32  # ring of pinned sparse buffers: frames decoded ahead + frames in flight
            ring_sp = [(U.SparseImage(), U.SparseImage()) for _ in range(K_)]
            pool = ThreadPoolExecutor(max_workers=args.jpeg_threads)
            ahead = 12

            def run_compressed(n, start):
                valid, q, bytes_up = 0, 0, 0
                futs = {}

                def submit(i):
                    k = pingpong(start + i, N_DISTINCT)
                    sl, sr = ring_sp[i % K_]
                    futs[i] = (pool.submit(sl.decode, enc[k][0]), pool.submit(sr.decode, enc[k][1]))
                for i in range(min(ahead, n)):
                    submit(i)
                for i in range(n):
                    fl, fr = futs.pop(i)
                    sl, sr = fl.result(), fr.result()
                    bytes_up += sl.nbytes + sr.nbytes
                    vo.enqueue_host_sparse(sl, sr, dt_frame)
                    q += 1
                    if q >= inflight:
                        valid += vo.collect().valid
                        q -= 1
                    if i + ahead < n:
                        submit(i + ahead)  # its ring slot held frame i + ahead - K_, collected long ago (K_ > ahead + inflight)
                while q:
                    valid += vo.collect().valid
                    q -= 1
                return valid, bytes_up
            run_compressed(3 * K_, 0)  # every pinned ring entry and every device staging slot allocated and touched
            comp_regions, comp_valid, comp_bytes = [], 0, 0
            n_comp = max(args.steps, 20)
            for r_ in range(min(args.regions, 5)):
                barrier()
                t0 = time.perf_counter()
                v_, b_ = run_compressed(n_comp, 3 * K_ + r_ * n_comp)
                barrier()
                comp_regions.append((time.perf_counter() - t0) * 1e3)
                comp_valid += v_
                comp_bytes += b_
            pool.shutdown()
            # the same with Huffman decoding ON THE GPU: one host thread hands the JPEG bytes to
            # uvo_stereo_enqueue_host_jpeg (marker walk + unstuffed copy of the scan on the host, nothing else)
            gpu_host_s = [0.0, 0]
            # more frames enqueued than lanes: the decode of a frame runs on its slot's own stream and overlaps the
            # lane work of the frames ahead of it
            comp_inflight = max(inflight, vo.max_in_flight() - 2)

            def run_compressed_gpu(n, start):
                valid, q = 0, 0
                for i in range(n):
                    k = pingpong(start + i, N_DISTINCT)
                    t_e = time.perf_counter()
                    vo.enqueue_host_jpeg(enc[k][0], enc[k][1], dt_frame)
                    gpu_host_s[0] += time.perf_counter() - t_e
                    gpu_host_s[1] += 1
                    q += 1
                    if q >= comp_inflight:
                        valid += vo.collect().valid
                        q -= 1
                while q:
                    valid += vo.collect().valid
                    q -= 1
                return valid
            run_compressed_gpu(48, 0)
            gpu_host_s[0], gpu_host_s[1] = 0.0, 0
            gpu_regions = []
            for r_ in range(min(args.regions, 5)):
                barrier()
                t0 = time.perf_counter()
                run_compressed_gpu(n_comp, 48 + r_ * n_comp)
                barrier()
                gpu_regions.append((time.perf_counter() - t0) * 1e3)
            # one long region of the same leg: the fill of a 14-deep pipeline (upload + 0.8 ms decode + the frame
            # itself) is a fifth of a 20-frame region
            n_long_c = max(10 * args.steps, 200)
            barrier()
            t0 = time.perf_counter()
            run_compressed_gpu(n_long_c, 48 + 5 * n_comp)
            barrier()
            gpu_long_ms = (time.perf_counter() - t0) * 1e3
            comp = {"gpu_inflight": comp_inflight, "gpu_region_ms": gpu_regions, "gpu_long": (n_long_c, gpu_long_ms), "gpu_entropy_frames": vo.gpu_entropy_frames,
                    "gpu_host_enqueue_us": 1e6 * gpu_host_s[0] / max(gpu_host_s[1], 1),
                    "gpu_h2d_bytes_per_step": float(np.mean([len(a) + len(b) for a, b in enc])) + 2 * 9500.0,"region_ms": comp_regions, "frames_per_region": n_comp, "valid": comp_valid,
                    "h2d_bytes_per_step": comp_bytes / float(n_comp * len(comp_regions)),
                    "jpeg_bytes_per_step": float(np.mean([len(a) + len(b) for a, b in enc])),
                    "host_threads": args.jpeg_threads}
        except ImportError:
            comp = None
    # ---- steady state: one long region each (pipeline fill / drain amortised), reported beside the contract's numbers
    n_long = max(10 * args.steps, 200)
    ms_long_dev, _ = timed_region(n_long, host=False)
    ms_long_host, _ = timed_region(n_long, host=True)
    # synchronous per-frame latency through uvo_stereo_frame (one frame in flight, H2D + D2H inside)
    n_sync = min(args.steps, 30)
    run_host_sync(2, pos[0])
    t0 = time.perf_counter()
    run_host_sync(n_sync, pos[0] + 2)
    sync_ms = (time.perf_counter() - t0) * 1e3 / n_sync
    pos[0] += n_sync + 2
    stop.set()
    sampler.join(timeout=3)

    # ---- per-kernel CUDA-event timing (roofline leg): same workload, events around every launch
    kern = {}
    stage = {}
    n3d = 0
    if rank == 0:
        ctx.kernel_timing(True)
        n_prof = min(max(args.steps, 20), 50)
        run_frames(n_prof, pos[0], inflight=1)   # one frame at a time: every kernel is timed running alone
        rep = ctx.kernel_report()
        ctx.kernel_timing(False)
        kern = {k: {"launches_per_frame": c / n_prof, "ms_per_frame": ms / n_prof, "us_per_launch": 1e3 * ms / c}
                for k, (c, ms) in rep.items()}
        L, R = dev_ring[(pos[0] + n_prof) % RING]
        r_last = vo.frame_device(L.data_ptr(), R.data_ptr(), pitch, dt_frame)
        n3d = int(r_last.n_3d)
        stage = vo.stage_ms()
    kl, _ = vo.last_keypoints(False)
    kr, _ = vo.last_keypoints(True)

    # ---- max over ranks (per region), then the median region
    comp_ms = float(np.median(comp["region_ms"])) if comp else 0.0
    comp_gpu_ms = float(np.median(comp["gpu_region_ms"])) if comp else 0.0
    comp_gpu_long_ms = comp["gpu_long"][1] if comp else 0.0
    all_ms, v = aggregate_over_ranks(dist if world > 1 else None,
                                     dev_regions + host_regions + [ms_long_dev, ms_long_host, comp_gpu_long_ms, comp_ms,
                                                                   comp_gpu_ms],
                                     [float(valid_dev), float(valid_host)], "cuda")
    comp_gpu_long_ms, comp_ms, comp_gpu_ms = all_ms[-3], all_ms[-2], all_ms[-1]
    R_ = args.regions
    dev_regions, host_regions = all_ms[:R_], all_ms[R_:2 * R_]
    ms_long_dev, ms_long_host = all_ms[2 * R_], all_ms[2 * R_ + 1]
    ms_dev, ms_host = float(np.median(dev_regions)), float(np.median(host_regions))
    total_frames = args.steps * world
    value = total_frames / (ms_dev * 1e-3)
    e2e = total_frames / (ms_host * 1e-3)

    if rank == 0:
        hbm, tf_burst, tf_sus, which = load_peaks()
        f64_peak, f64_src = fp64_peak()
        n_kp = (len(kl) + len(kr)) / 2.0
        P = W * H
        algo = kernel_models(P, n_kp, n3d, p.iterations_count)
        # DRAM bytes per launch of each kernel from the committed `ncu --set full` capture of this same command
        # (profiles/ncu_traffic.json, written by tools/ncu_summary.py); null when the kernel was not captured
        traffic = {}
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch", {})

        def roof_of(name, kt):
            if name not in algo:
                return None
            bound, per_launch = algo[name]
            sec = kt["us_per_launch"] * 1e-6
            if bound == "hbm":
                ach, peak, unit, src = per_launch / sec / 1e9, hbm, "GB/s", which
            elif bound == "tensor":
                ach, peak, unit, src = per_launch / sec / 1e12, tf_sus, "TFLOP/s", which + " (sustained bf16)"
            else:
                ach, peak, unit, src = per_launch / sec / 1e12, f64_peak, "TFLOP/s", f64_src
            r = {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                 "traffic": traffic.get(name.split("[")[0]), "peak_source": src, "us_per_launch": kt["us_per_launch"],
                 ("algorithmic_bytes_per_launch" if bound == "hbm" else "algorithmic_flops_per_launch"): per_launch}
            if bound == "tensor":
                r["note"] = ("tcgen05 kind::tf32 (nominal dense peak is half of bf16's); peak quoted is the measured "
                             "bf16 figure as the profiling recipe prescribes")
            if bound == "fp64":
                r["note"] = ("FP64-issue roofline: sustained DFMA rate of the chip; the kernel is one dependent chain "
                             "per hypothesis (Jacobi rotations: sqrt, sqrt, reciprocal), so it is bound by the 8-cycle "
                             "DFMA latency, not by the issue rate")
            return r

        # the dominant kernel: top by time per frame over ALL kernels (each timed running alone)
        top_overall = max(kern.items(), key=lambda kv: kv[1]["ms_per_frame"])[0] if kern else None
        roof = roof_of(top_overall, kern[top_overall]) if top_overall else None
        # beside it, the top WIDE kernel: the pose kernels that lead by duration run on 1-32 blocks and overlap the other
        # frames in flight; the kernel that leads among those whose grid fills the chip is what bounds the frame rate
        wide = {n: kt for n, kt in kern.items() if n in algo and algo[n][0] != "fp64"}
        top_wide = max(wide.items(), key=lambda kv: kv[1]["ms_per_frame"])[0] if wide else None
        roof_wide = roof_of(top_wide, kern[top_wide]) if top_wide else None
        rooflines = {}
        for name, kt in kern.items():
            r = roof_of(name, kt)
            if r:
                rooflines[name] = {"bound": r["bound"], "achieved": r["achieved"], "unit": r["unit"], "frac": r["frac"],
                                   "us_per_launch": kt["us_per_launch"]}
        # CPU baseline: bounded sample of the same workload on the host cores (oracle port)
        cpu = None
        if not args.no_cpu and world == 1:  # the CPU baseline is reported by the single-GPU run only
            n_cpu = 4
            spf, _ = cpu_frames(seq, thr, n_cpu, params=p)
            cpu = {"value": 1.0 / spf, "unit": "frames/s", "cores": min(16, os.cpu_count() or 1), "kind": "port",
                   "sample": f"{n_cpu} consecutive stereo frames (after the init frame) of this run's sequence through "
                             "the oracle port of the OpenCV CPU path; SURF restated, not OpenCV",
                   "host_cpus": os.cpu_count(),
                   "stage_ms": getattr(cpu_frames, "last_stage_ms", None),
                   "stage_ms_cv2": getattr(cpu_frames, "last_stage_ms_cv2", None)}
        frames_timed = args.steps * args.regions
        line = {
            "metric": "stereo_uvo_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32/f32/f64",
            "data": "synthetic",
            "config": config_dict(cfg, args.config, thr),
            "run": {"keypoints_per_image": n_kp, "sequences_per_gpu": 1, "frames_in_flight": inflight,
                    "cuda_graphs": bool(args.graphs)},
            "n_sequences": world,
            # every number above is the MEDIAN of `regions` consecutive timed regions of exactly `steps` frames, each
            # bracketed by barrier + synchronize (max over ranks per region); a 20-frame region is 7 ms long against a
            # 1 ms pipeline fill, so single regions scatter by several percent
            "timing": {"regions": args.regions, "region_ms": dev_regions, "spread": (max(dev_regions) - min(dev_regions)) /
                       ms_dev, "e2e_region_ms": host_regions,
                       "e2e_spread": (max(host_regions) - min(host_regions)) / ms_host,
                       "steady_state": {"frames": n_long, "value": n_long * world / (ms_long_dev * 1e-3),
                                        "e2e": n_long * world / (ms_long_host * 1e-3)}},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": 2 * 3 * P,
                    "d2h_bytes_per_step": int(C.sizeof(U.StereoResult)),
                    "ms_per_step": ms_host / args.steps,
                    "api": f"uvo_stereo_enqueue_host + uvo_stereo_collect (pinned host images, {inflight} frames in "
                           "flight; H2D of both images and D2H of the result record inside the timed region)",
                    "sync_frame_latency_ms": sync_ms},
            # the same metric with the node's real input: compressed images (JPEG quality 90, 4:2:0) in host memory,
            # Huffman decoding on host threads inside the timed region, sparse coefficients over PCIe
            "e2e_compressed": None if not comp else {
                "value": comp["frames_per_region"] * world / (comp_ms * 1e-3), "unit": "frames/s",
                "h2d_bytes_per_step": comp["h2d_bytes_per_step"], "d2h_bytes_per_step": int(C.sizeof(U.StereoResult)),
                "jpeg_bytes_per_step": comp["jpeg_bytes_per_step"], "host_threads_per_gpu": comp["host_threads"],
                "region_ms": comp["region_ms"],
                "api": "uvo_jpeg_entropy_decode_sparse on host threads + uvo_stereo_enqueue_host_sparse + "
                       "uvo_stereo_collect; wall clock, barrier + synchronize on both sides",
                # Huffman decoding on the GPU (k_jpeg_huff): one host thread, JPEG bytes in, scan bytes over PCIe
                "gpu_entropy": {"value": comp["frames_per_region"] * world / (comp_gpu_ms * 1e-3), "unit": "frames/s",
                                "h2d_bytes_per_step": comp["gpu_h2d_bytes_per_step"], "host_threads_per_gpu": 1,
                                "frames_in_flight": comp["gpu_inflight"],
                                "steady_state": {"frames": comp["gpu_long"][0],
                                                 "value": comp["gpu_long"][0] * world / (comp_gpu_long_ms * 1e-3)},
                                "region_ms": comp["gpu_region_ms"], "frames_on_gpu_decoder": comp["gpu_entropy_frames"],
                                "host_enqueue_us_per_frame": comp["gpu_host_enqueue_us"],
                                "api": "uvo_stereo_enqueue_host_jpeg + uvo_stereo_collect"}},
            "host_enqueue_us_per_frame": 1e6 * host_enqueue_s[0] / max(host_enqueue_s[1], 1),
            "host_affinity": affinity,
            # kernels the library launched inside the timed regions of `value` (graph-replayed kernels counted one by
            # one), per region of `steps` frames; host-side launch calls are graph launches + direct launches
            "gpu_launches": int(round(launches / float(args.regions))),
            "launches_per_frame": launches / float(frames_timed),
            "graph_launches_per_frame": graph_launches / float(frames_timed),
            "valid_frames": {"device": int(v[0]), "host": int(v[1]), "of": frames_timed * world},
            "clocks": summarise_clocks(samples),
            "roofline": roof, "roofline_top_wide_kernel": roof_wide, "rooflines_all": rooflines,
            "top_kernel_by_time": top_overall, "cpu_baseline": cpu,
            "fp64_peak_tflops": f64_peak,
            "stage_ms": stage, "kernels": kern,
        }
        print(json.dumps(line))
    vo.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="B", choices=sorted(CONFIGS), help="BASELINE.json config: B = configs[1] "
                    "(1280x1024, the headline), E = configs[4] (2448x2048, one sequence per GPU)")
    ap.add_argument("--regions", type=int, default=REGIONS, help="consecutive timed regions of --steps frames (median)")
    ap.add_argument("--threshold", type=int, default=0, help="SURF min_hessian (0: bisect for ~4096 keypoints)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--inflight", type=int, default=0, help="frames kept in flight per sequence (0: the library's lanes)")
    ap.add_argument("--graphs", type=int, default=1, help="0: launch every kernel directly instead of replaying graphs")
    ap.add_argument("--jpeg-threads", type=int, default=8, help="host threads of the compressed-input leg (0: skip it)")
    args = ap.parse_args()
    args.regions = max(args.regions, 1)
    if args.impl == "reference":
        if args.steps > 20:
            args.steps = 20  # bounded sample: ~2 s of CPU work per frame
        args.warmup = min(args.warmup, 5)
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
