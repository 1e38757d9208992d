import io, sys, ctypes as C, numpy as np, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from PIL import Image
from spatialaudiogen_b200 import _lib as L
lib=L.lib()
rng=np.random.RandomState(11)
def pic(h,w): return np.clip(np.kron(rng.randint(0,256,((h+5)//6,(w+5)//6,3)),np.ones((6,6,1)))[:h,:w]+rng.randn(h,w,3)*10,0,255).astype(np.uint8)
bases=[]
for h,w,kw in ((40,56,dict(quality=80,subsampling=2)),(33,49,dict(quality=60,subsampling=1)),(24,24,dict(quality=95,subsampling=0)),(48,64,dict(quality=70,subsampling=2,restart_marker_blocks=2)),(30,30,dict(quality=50))):
    b=io.BytesIO(); Image.fromarray(pic(h,w)).save(b,'JPEG',**kw); bases.append((h,w,b.getvalue()))
ok=err=0; t0=time.time()
for it in range(6000):
    h,w,data=bases[it%len(bases)]
    d=bytearray(data)
    mode=it%4
    n=rng.randint(1,4)
    for _ in range(n):
        if mode==0: pos=rng.randint(2,min(len(d),700))            # headers / tables
        elif mode==1: pos=rng.randint(len(d)//2,len(d))           # scan data
        else: pos=rng.randint(2,len(d))
        if mode==3 and rng.rand()<0.5: d=d[:pos]; break           # truncation
        d[pos]=rng.randint(256)
    d=bytes(d)
    cap=3*((h+15)//16*16)*((w+15)//16*16)
    # geometry may change with header mutations: size the buffer from sag_jpeg_info when it parses
    v=[C.c_int() for _ in range(5)]
    rc=lib.sag_jpeg_info(d,len(d),*[C.byref(x) for x in v])
    if rc!=0: err+=1; continue
    W,H=v[0].value,v[1].value
    if W*H>4_000_000: continue
    cap=3*((H+15)//16*16)*((W+15)//16*16)
    out=np.zeros(cap,np.int16); bw=(C.c_int*3)(); bh=(C.c_int*3)()
    rc1=lib.sag_jpeg_coefficients(d,len(d),out.ctypes.data,out.size,bw,bh,None)
    out2=np.zeros(cap,np.int16); r=C.c_int()
    rc2=lib.sag_jpeg_coefficients_parallel(d,len(d),int(rng.choice([1,7,64])),out2.ctypes.data,out2.size,C.byref(r))
    if rc1==0 and rc2==0:
        ok+=1
    else: err+=1
print('mutated files',6000,'decoded',ok,'rejected',err,'%.1fs'%(time.time()-t0),'(no crash)')
