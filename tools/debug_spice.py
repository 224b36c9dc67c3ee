import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, warnings
warnings.simplefilter("ignore")
from euispice_coreg_b200._synth.spice import make_spice_case, small_spice_spec
from euispice_coreg_b200._compat import fits_lite
from euispice_coreg_b200._compat.wcs import TanWcs
from euispice_coreg_b200 import _ext
from euispice_coreg_b200.hdrshift import AlignmentSpice
from euispice_coreg_b200.synras import SPICEComposedMapBuilder
from oracle.hpc import HpcSearch, shift_header
from oracle import wcs_tan
from oracle.synras import spice_l2_image
d='/tmp/spice_dbg'
p_spice, imagers, spec = make_spice_case(d, small_spice_spec(nbin2=8, pxbeg2=192), tag="toy")
synras = SPICEComposedMapBuilder(p_spice, imagers, threshold_time=100.0).process(folder_path_output=d, basename_output="synras2.fits", print_filename=False, return_synras_name=True)
lags = dict(lag_crval1=np.array([-8.0]), lag_crval2=np.array([12.0]), lag_cdelt1=[0], lag_cdelt2=[0], lag_crota=[0])
a = AlignmentSpice(synras, p_spice, parallelism=True, small_fov_window=0, large_fov_window=-1, **lags)
gpu = a.align_using_helioprojective(return_type="corr")
ld=lambda q:(fits_lite.open(q)[0].data, dict(fits_lite.open(q)[0].header.items()))
d4,h4=ld(p_spice); img,hdr=spice_l2_image(d4,h4); dl,hl=ld(synras)
s=HpcSearch(dl,hl,img,hdr,**lags); ref=s.cube()
print('gpu',gpu.ravel(),'ref',ref.ravel(), 'nvalid', a.nvalid.ravel())
r=a.engine.ref.cpu().numpy()
print('K2 ref equal:', np.array_equal(r, s.data_large, equal_nan=True), np.nanmax(np.abs(r-s.data_large)), np.isnan(r).sum(), np.isnan(s.data_large).sum())
h=dict(s.hdr_small); shift_header(h, s.refs, s.refs.lag_crval1[0], s.refs.lag_crval2[0], 0.0,0.0,0.0)
lng,lat=s.world()
xo,yo=wcs_tan.WcsTan(h).world_to_pixel(lng,lat)
x,y=_ext.tan_world2pix(TanWcs.from_header(h), torch.from_numpy(lng).cuda(), torch.from_numpy(lat).cuda())
print('w2p diff', np.abs(x.cpu().numpy()-xo).max(), np.abs(y.cpu().numpy()-yo).max())
interp=s.reprojected(s.refs.lag_crval1[0], s.refs.lag_crval2[0],0.0,0.0,0.0)
valid=np.isfinite(interp)&np.isfinite(s.data_large)
print('oracle nvalid', valid.sum())
print('hdr_small product', {k:a.hdr_small[k] for k in ('CRVAL1','CRVAL2','CDELT1','CDELT2','CUNIT1','CROTA') }, a.hdr_small.get('PC1_1'), a.hdr_small.get('PC1_2'), a.hdr_small.get('PC2_1'))
print('hdr oracle', {k:hdr[k] for k in ('CRVAL1','CRVAL2','CDELT1','CDELT2','CUNIT1','CROTA','PC1_1','PC1_2','PC2_1')})
print('lag arrays product', a.lag_crval1, a.lag_crval2, a.unit_lag, 'oracle', s.refs.lag_crval1, s.refs.lag_crval2)
