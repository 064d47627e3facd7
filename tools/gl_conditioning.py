"""How well conditioned is 32-iteration Griffin-Lim on the smoke() clips?  CPU only (oracle code, torch fp32 / fp64).

VERDICT r1 #4 asked where the dB between the GPU path and the CPU oracle go (49.9 dB on one of the two smoke clips while
the other sits at 89 dB).  tests/test_gpu_parity.py::test_exact_math_attribution shows that no approximation on the GPU
path explains it (all-exact mode: 51.0 dB).  This script shows that the ORACLE disagrees with ITSELF by the same amount on
that clip when its input magnitudes move by 1e-6 relative -- ten times less than the parity budget of the stages that
produce them: a bin whose rebuilt value nearly cancels flips its direction under the unit-modulus projection, and the flip
is worth ~50 dB.  Output committed under profiles/r2_gl_conditioning.txt.
"""
import sys, json, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dsp, metrics, model as omodel, pipeline as opipe, synth
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'weights_good.npz'))
cfg = json.loads(bytes(z["__config__"]).decode())
sd = {k: torch.from_numpy(z[k].copy()) for k in z.files if k != "__config__"}
noisy,_ = synth.make_batch(2,16000,16000)
T=1+16000//512
init = synth.gl_init_angles((2,513,T),seed=7)
orc = omodel.GRUUNet2Oracle(sd,cfg)
ref = opipe.denoise_batch(noisy, orc, 1024,512,64,16000,32,0.99,init)
lin = ref['lin_mag']; peak = ref['peak']
w32 = dsp.griffinlim(lin,1024,512,32,0.99,init)
# float64 GL of the same magnitudes / init
import torchaudio.functional as TAF
def gl64(lin, init):
    lin=lin.double(); 
    return dsp.griffinlim(lin,1024,512,32,0.99,init.to(torch.complex128))
try:
    w64 = gl64(lin,init)
    print('fp32 vs fp64 GL SI-SDR', metrics.si_sdr(w32, w64.float()).tolist())
except Exception as e:
    print('gl64 failed', e)
# perturb lin by 1 ulp-level relative noise
g=torch.Generator().manual_seed(0)
for eps in (1e-7, 3e-7, 1e-6):
    linp = lin*(1+eps*torch.randn(lin.shape,generator=g))
    wp = dsp.griffinlim(linp,1024,512,32,0.99,init)
    print('eps',eps,'SI-SDR', [round(float(v),1) for v in metrics.si_sdr(wp,w32)])
# perturb init angles by 1e-7
for eps in (1e-7,):
    ip = init*(1+eps*torch.randn(init.shape,generator=g))
    wp = dsp.griffinlim(lin,1024,512,32,0.99,ip)
    print('init eps',eps,'SI-SDR', [round(float(v),1) for v in metrics.si_sdr(wp,w32)])
print('---- larger perturbations of lin_mag')
for eps in (1e-6, 3e-6, 1e-5, 3e-5):
    for rep in range(3):
        linp = lin*(1+eps*torch.randn(lin.shape,generator=g))
        wp = dsp.griffinlim(linp,1024,512,32,0.99,init)
        print('eps',eps,'SI-SDR', [round(float(v),1) for v in metrics.si_sdr(wp,w32)])
