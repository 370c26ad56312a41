"""Mutation fuzzer for the host half of the JPEG decode (uvo_jpeg_info, uvo_jpeg_entropy_decode,
uvo_jpeg_entropy_decode_sparse, and uvo_jpeg_gpu_plan -- the host half of the GPU Huffman route): valid streams (all sub-samplings, with and without restart intervals, grayscale) with
random byte edits, truncations, header edits and insertions; every call must return a status code, and the sparse
tables must stay inside their bounds.  The streams arrive from the network (ROS CompressedImage messages), so the parser
must not trust them.

    python tools/jpeg_fuzz.py [seed] [iterations]
Under AddressSanitizer: build a copy of the library with
    make -C <copy>/ergo_uvo_b200/csrc EXTRA="-Xcompiler -fsanitize=address,-fno-omit-frame-pointer -g"
and run with UVO_ROOT=<copy> ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0
    LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so.6)"."""
import sys; ROOT = __import__('os').environ.get('UVO_ROOT', __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))), 'tests'))
import numpy as np, cv2, ctypes as C
import ergo_uvo_b200 as U
from ergo_uvo_b200 import _lib as L
from conftest import noise_image
lib=L.load()
lib.uvo_jpeg_gpu_staging_bytes.restype=C.c_size_t
rs=np.random.RandomState(int(sys.argv[1]) if len(sys.argv)>1 else 0)
img=noise_image(67,93,seed=1,channels=3)
seeds=[]
for sf in ("444","420","422","411"):
    for rst in (0,2):
        ok,enc=cv2.imencode(".jpg",img,[cv2.IMWRITE_JPEG_QUALITY,70,cv2.IMWRITE_JPEG_SAMPLING_FACTOR,getattr(cv2,"IMWRITE_JPEG_SAMPLING_FACTOR_"+sf),cv2.IMWRITE_JPEG_RST_INTERVAL,rst]); seeds.append(enc.tobytes())
ok,enc=cv2.imencode(".jpg",img[:,:,0].copy()); seeds.append(enc.tobytes())
codes={}
N=int(sys.argv[2]) if len(sys.argv)>2 else 20000
for it in range(N):
    d=bytearray(seeds[rs.randint(len(seeds))])
    kind=rs.randint(4)
    if kind==0:
        for _ in range(rs.randint(1,6)): d[rs.randint(len(d))]=rs.randint(256)
    elif kind==1:
        d=d[:rs.randint(2,len(d))]
    elif kind==2:
        p=rs.randint(2,min(len(d),700)); d[p]=rs.randint(256)   # header area
    else:
        p=rs.randint(len(d)); d[p:p]=bytes(rs.randint(0,256,rs.randint(1,8)).astype(np.uint8))
    buf=np.frombuffer(bytes(d),np.uint8)
    lay=L.JpegLayout()
    rc=lib.uvo_jpeg_info(buf.ctypes.data_as(C.c_void_p),C.c_size_t(len(buf)),C.byref(lay))
    codes[("info",rc)]=codes.get(("info",rc),0)+1
    # the host half of the GPU Huffman route (marker walk, table plan, unstuffed scan copy) on the same bytes: the
    # staging buffer is allocated at exactly the size the library asks for, so an overrun is the sanitizer's to find
    stg=np.empty(int(lib.uvo_jpeg_gpu_staging_bytes(C.c_size_t(len(buf)))),np.uint8)
    q=C.c_int(0); up=C.c_size_t(0); bits=C.c_uint32(0); lay2=L.JpegLayout()
    rc2=lib.uvo_jpeg_gpu_plan(buf.ctypes.data_as(C.c_void_p),C.c_size_t(len(buf)),stg.ctypes.data_as(C.c_void_p),C.c_size_t(len(stg)),C.byref(q),C.byref(up),C.byref(bits),C.byref(lay2))
    codes[("plan",rc2,q.value)]=codes.get(("plan",rc2,q.value),0)+1
    if rc2==0 and q.value:
        assert up.value<=len(stg) and (bits.value+7)//8+16<=up.value and rc==0 and lay2.coeff_total==lay.coeff_total
    if rc2==0: assert rc==0   # what the GPU route accepts, the host parser accepts
    if rc==0:
        tot=int(lay.coeff_total)
        if tot<=0 or tot>64*1000000: 
            codes[("huge",0)]=codes.get(("huge",0),0)+1; continue
        coef=np.empty(tot,np.int16)
        rc=lib.uvo_jpeg_entropy_decode(buf.ctypes.data_as(C.c_void_p),C.c_size_t(len(buf)),coef.ctypes.data_as(C.c_void_p),C.c_size_t(tot),C.byref(lay))
        codes[("dense",rc)]=codes.get(("dense",rc),0)+1
        ent=np.empty(tot,np.uint32); first=np.empty(tot//64,np.uint32); cnt=np.empty(tot//64,np.uint8); n=C.c_size_t(0)
        rc=lib.uvo_jpeg_entropy_decode_sparse(buf.ctypes.data_as(C.c_void_p),C.c_size_t(len(buf)),ent.ctypes.data_as(C.c_void_p),C.c_size_t(tot),first.ctypes.data_as(C.c_void_p),cnt.ctypes.data_as(C.c_void_p),C.byref(n),C.byref(lay))
        codes[("sparse",rc)]=codes.get(("sparse",rc),0)+1
        if rc==0:
            assert n.value<=tot and (first[cnt>0].astype(np.int64)+cnt[cnt>0]<=n.value).all()
print(codes)
