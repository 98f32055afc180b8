"""Host-side PNG decode rate of the input stage: zlib inflate vs the library's own inflate vs the
full decode (inflate + un-filter), on 1241x376 frames written with the Sub filter (OpenCV's default) and with
libpng's adaptive filters (Paeth-heavy).  CPU only.  usage: python tools/png_bench.py"""
import sys, time, zlib, struct, numpy as np, cv2
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from importlib import import_module
io=import_module('r7020e-visual-odometry_b200.io'); synth=import_module('r7020e-visual-odometry_b200.synth')
l,r=synth.shift_stream(2,seed=5)
L,R=synth.plane_world([np.eye(4)], seed=3)
def tm(f,n=15,rep=7):
    f(); best=1e9
    for _ in range(rep):
        t=time.perf_counter()
        for _ in range(n): f()
        best=min(best,(time.perf_counter()-t)/n*1e3)
    return best
for iname,im in (('texture',l[0]),('plane_world',L[0])):
    for name,params in [('cv2 default',[]),('libpng adaptive l6',[cv2.IMWRITE_PNG_COMPRESSION,6, cv2.IMWRITE_PNG_STRATEGY, cv2.IMWRITE_PNG_STRATEGY_DEFAULT])]:
        ok,buf=cv2.imencode('.png',im,params); b=buf.tobytes()
        pos=8; idat=b''
        while pos<len(b):
            ln=struct.unpack('>I',b[pos:pos+4])[0]; typ=b[pos+4:pos+8]
            if typ==b'IDAT': idat+=b[pos+8:pos+8+ln]
            pos+=12+ln
        raw=zlib.decompress(idat)
        print(f"{iname:12s} {name:20s} zlib {tm(lambda: zlib.decompress(idat)):.2f}  fast {tm(lambda: io.inflate_zlib(idat,len(raw))):.2f}  full decode {tm(lambda: io.png_decode(b)):.2f} ms")
