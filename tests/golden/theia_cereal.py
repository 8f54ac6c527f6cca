"""Reader for TheiaSfM reconstructions serialised with cereal's PortableBinary archive (the format of the reference's
own fixtures data/sfm/fountain11.bin and data/sfm/gt_fountain11.bin; writer: io/reconstruction_writer.cc:52-70).
Layouts followed: Reconstruction (reconstruction.h:159-167), View (view.h:92-94), Camera v0/v1 (camera.h:207-249),
CameraIntrinsicsModel / PinholeCameraModel (camera_intrinsics_model.h:216-218, pinhole_camera_model.h:169-178),
CameraIntrinsicsPrior v0-v4 (camera_intrinsics_prior.h:102-136), Track (track.h:81-83), Eigen matrices
(io/eigen_serializable.h).  cereal rules used: 1-byte endianness flag, a u32 class version the first time a versioned
type appears, u64 container sizes, polymorphic shared_ptr = polymorphic id (+name first time) + pointer id (+data first time).
Used only by tests/golden/make_fountain_fixture.py (fixture loaders into the IR)."""
import struct, sys
import numpy as np

class R:
    def __init__(self, b): self.b=b; self.o=0; self.seen=set()
    def u8(self): v=self.b[self.o]; self.o+=1; return v
    def u32(self): v=struct.unpack_from('<I', self.b, self.o)[0]; self.o+=4; return v
    def i32(self): v=struct.unpack_from('<i', self.b, self.o)[0]; self.o+=4; return v
    def u64(self): v=struct.unpack_from('<Q', self.b, self.o)[0]; self.o+=8; return v
    def f64(self, n=1):
        v=np.frombuffer(self.b, '<f8', n, self.o).copy(); self.o+=8*n; return v
    def s(self): n=self.u64(); v=self.b[self.o:self.o+n].decode(); self.o+=n; return v
    def ver(self, name):
        if name in self.seen: return self.vers[name]
        self.seen.add(name); v=self.u32(); self.vers=getattr(self,'vers',{}); self.vers[name]=v; return v

def prior(r, n):
    r.ver('Prior%d'%n)
    is_set=r.u8(); val=r.f64(n); return (is_set, val)

def cam_prior(r):
    v=r.ver('CameraIntrinsicsPrior')
    out={'version':v}
    if v>=4:
        out['w']=r.i32(); out['h']=r.i32(); out['model']=r.s()
        out['focal']=prior(r,1); out['pp']=prior(r,2); out['aspect']=prior(r,1); out['skew']=prior(r,1)
        out['rad']=prior(r,4); out['tan']=prior(r,2); out['pos']=prior(r,3); out['ori']=prior(r,3)
        out['lat']=prior(r,1); out['lon']=prior(r,1); out['alt']=prior(r,1)
    elif v==3:
        out['w']=r.i32(); out['h']=r.i32(); out['model']=r.s()
        out['focal']=prior(r,1); out['aspect']=prior(r,1); out['skew']=prior(r,1)
        out['rad']=prior(r,4); out['tan']=prior(r,2); out['pos']=prior(r,3); out['ori']=prior(r,3)
        out['lat']=prior(r,1); out['lon']=prior(r,1); out['alt']=prior(r,1)
    elif v==2:
        out['w']=r.i32(); out['h']=r.i32()
        out['focal']=prior(r,1); out['aspect']=prior(r,1); out['skew']=prior(r,1)
        out['rad']=prior(r,2); out['tan']=prior(r,2); out['pos']=prior(r,3); out['ori']=prior(r,3)
        out['lat']=prior(r,1); out['lon']=prior(r,1); out['alt']=prior(r,1)
    else:
        if v>=1: out['w']=r.i32(); out['h']=r.i32()
        out['focal']=prior(r,1); out['ppx']=prior(r,1); out['ppy']=prior(r,1); out['aspect']=prior(r,1); out['skew']=prior(r,1)
        out['rd1']=prior(r,1); out['rd2']=prior(r,1)
    return out

def camera(r):
    v=r.ver('Camera')
    if v==0:
        p=r.f64(13); img=(r.i32(), r.i32())
        return {'ext':p[:6], 'intr':p[6:], 'model':0, 'img':img, 'version':0}
    ext=r.f64(6)
    MSB=0x80000000
    pid=r.u32()
    names=getattr(r,'poly_names',{}); r.poly_names=names
    if pid==0: raise ValueError('null intrinsics')
    if pid & MSB:
        names[pid & ~MSB]=r.s()
    pname=names[pid & ~MSB]
    sid=r.u32()
    ptrs=getattr(r,'ptrs',{}); r.ptrs=ptrs
    if sid & MSB:
        mv=r.ver(pname)
        if 'PinholeCameraModel' in pname and 'Radial' not in pname:
            if mv>0:
                r.ver('CameraIntrinsicsModel'); n=r.u64(); params=r.f64(n)
            else:
                params=r.f64(7)
        else:
            r.ver('CameraIntrinsicsModel'); n=r.u64(); params=r.f64(n)
        ptrs[sid & ~MSB]={'params':params,'type':pname}
    obj=ptrs[sid & ~MSB]
    img=(r.i32(), r.i32())
    return {'ext':ext,'intr':obj['params'],'intr_id':sid & ~MSB,'type':obj['type'],'img':img,'version':v}

def eigen(r, dtype='<f8'):
    rows=r.i32(); cols=r.i32(); n=rows*cols
    it=np.dtype(dtype).itemsize
    v=np.frombuffer(r.b, dtype, n, r.o).copy(); r.o+=n*it; return v

def view(r):
    v=r.ver('View')
    out={'name':r.s(), 'est':r.u8(), 'camera':camera(r), 'prior':cam_prior(r)}
    n=r.u64(); feats={}
    for _ in range(n):
        t=r.u32(); feats[t]=eigen(r)
    out['features']=feats
    return out

def track(r):
    v=r.ver('Track')
    est=r.u8(); n=r.u64(); views=[r.u32() for _ in range(n)]
    pt=eigen(r); color=eigen(r,'u1')
    return {'est':est,'views':views,'pt':pt,'color':color}

def reconstruction(b):
    r=R(b)
    assert r.u8()==1
    r.ver('Reconstruction')
    out={'next_track':r.u32(), 'next_view':r.u32()}
    n=r.u64(); out['name2id']={}
    for _ in range(n):
        k=r.s(); out['name2id'][k]=r.u32()
    n=r.u64(); out['views']={}
    for _ in range(n):
        k=r.u32(); out['views'][k]=view(r)
    n=r.u64(); out['tracks']={}
    for _ in range(n):
        k=r.u32(); out['tracks'][k]=track(r)
    n=r.u64(); out['view2group']={}
    for _ in range(n):
        k=r.u32(); out['view2group'][k]=r.u32()
    n=r.u64(); out['groups']={}
    for _ in range(n):
        k=r.u32(); m=r.u64(); out['groups'][k]=[r.u32() for _ in range(m)]
    out['consumed']=r.o; out['total']=len(b); out['vers']=r.vers
    return out

