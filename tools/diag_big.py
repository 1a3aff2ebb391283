import sys, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import spectral_ref as sref
from sdr_iq_visualizer_b200 import spectral as sp
for nfft,hop,kind in [(262144,262144,'rect'),(1048576,524288,'hann'),(131072,65536,'hann'),(262144,131072,'hann')]:
    F=2; L=nfft+hop*(F-1)+7
    x=sref.synth_iq(L, seed=nfft%1000+3).astype(np.complex64)
    pl=sp.SpectralPlan(nfft,hop,kind)
    r=pl.stft(x, spectrum=True, db_rows=True)
    w=sref.window(kind,nfft)
    if kind!='rect': w=w.astype(np.float32).astype(np.float64)
    X=sref.shift_bins(np.fft.fft(sref.frames(x.astype(np.complex128),nfft,hop)*w,axis=1))
    P=np.abs(X)**2; rms=np.sqrt(P.mean())
    err=np.abs(r.spectrum-X)
    ratio=err/(6e-6*rms+1e-6*np.abs(X))
    i=np.unravel_index(np.argmax(ratio),ratio.shape)
    print(nfft,kind,'rms %.1f maxerr %.3e (%.2e rms) worst ratio %.2f at %s |X|=%.1f err=%.3e ; err rms %.2e rms; peak|X| %.0f err@peak %.3e'%(rms,err.max(),err.max()/rms,ratio.max(),i,abs(X[i]),err[i], np.sqrt((err**2).mean())/rms, np.abs(X).max(), err.flat[np.argmax(np.abs(X))]))
    pl.close()
