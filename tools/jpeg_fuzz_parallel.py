import io, sys, ctypes as C, numpy as np, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from PIL import Image
from spatialaudiogen_b200 import _lib as L
lib=L.lib()
rng=np.random.RandomState(int(sys.argv[1]) if len(sys.argv)>1 else 0)
N=int(sys.argv[2]) if len(sys.argv)>2 else 300
bad=0; maxr=0; t0=time.time()
for it in range(N):
    h,w=int(rng.randint(1,200)),int(rng.randint(2,260))
    kind=rng.randint(4)
    if kind==0: img=rng.randint(0,256,(h,w,3))
    elif kind==1:
        y,x=np.mgrid[0:h,0:w]; img=np.stack([127+100*np.sin(x/17.+y/9.+it),127+90*np.cos(x/5.-y/23.),127+80*np.sin(x/7.)*np.cos(y/3.)],-1)+rng.randn(h,w,3)*rng.uniform(0,40)
    elif kind==2: img=np.full((h,w,3),rng.randint(256))+rng.randn(h,w,3)*rng.uniform(0,3)
    else: img=np.kron(rng.randint(0,256,((h+7)//8,(w+7)//8,3)),np.ones((8,8,1)))[:h,:w]
    img=np.clip(img,0,255).astype(np.uint8)
    kw=dict(quality=int(rng.choice([1,5,20,50,75,90,97,100])))
    gray=rng.rand()<0.15
    if not gray: kw['subsampling']=int(rng.randint(3))
    if rng.rand()<0.3: kw['optimize']=True
    if rng.rand()<0.3: kw['restart_marker_blocks']=int(rng.randint(1,9))
    elif rng.rand()<0.2: kw['restart_marker_rows']=int(rng.randint(1,4))
    b=io.BytesIO()
    try: Image.fromarray(img[:,:,0] if gray else img).save(b,'JPEG',**kw)
    except Exception as e: print('save failed',kw,e); continue
    data=b.getvalue()
    cap=3*((h+15)//16*16)*((w+15)//16*16)
    ref=np.zeros(cap,np.int16); bw=(C.c_int*3)(); bh=(C.c_int*3)()
    rc=lib.sag_jpeg_coefficients(data,len(data),ref.ctypes.data,ref.size,bw,bh,None)
    if rc: print('serial rc',rc,lib.sag_last_error(),kw,h,w); bad+=1; continue
    n=sum(bw[c]*bh[c]*64 for c in range(3))
    for sub in (32,64,256):
        assert lib.sag_jpeg_set_option(None,b'sub_bytes',sub)==0
        for nt in (1,5,64):
            out=np.full(cap,-7,np.int16); r=C.c_int()
            rc=lib.sag_jpeg_coefficients_parallel(data,len(data),nt,out.ctypes.data,out.size,C.byref(r))
            maxr=max(maxr,r.value)
            if rc or not np.array_equal(out[:n],ref[:n]):
                bad+=1; print('MISMATCH',it,h,w,kw,gray,sub,nt,rc,r.value)
lib.sag_jpeg_set_option(None,b'sub_bytes',256)
print('images',N,'bad',bad,'max rounds',maxr,'%.1fs'%(time.time()-t0))
