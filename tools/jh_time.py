import numpy as np, cv2, time, sys
sys.path.insert(0, ".")
import ergo_uvo_b200 as U
from tools import synth
seq = synth.StereoSequence(1280, 1024, n_frames=1, seed=1300, tex_size=2048)
ok, enc = cv2.imencode(".jpg", seq.frames[0][0], [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420])
ctx = U.Context(0)
d = enc.tobytes()
for r in range(3): ctx.jpeg_decode(d)
ctx.kernel_timing(True)
for r in range(10): img = ctx.jpeg_decode(d)
print(len(d), ctx.jpeg_gpu_entropy(), {k: round(1e3*v[1]/v[0],1) for k,v in ctx.kernel_report().items()})
print(np.array_equal(img, cv2.imdecode(enc, cv2.IMREAD_UNCHANGED)))

import ctypes as C
st = (C.c_int64 * 24)()
n = C.c_int(0)
ctx.lib.uvo_jpeg_gpu_entropy_stamps(ctx.h, st, C.byref(n))
t = [st[i] for i in range(n.value)]
print("phase us:", [round((b - a) / 1e3, 1) for a, b in zip(t, t[1:])], "total", round((t[-1] - t[0]) / 1e3, 1))
