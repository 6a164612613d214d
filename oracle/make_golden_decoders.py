"""TEST INFRASTRUCTURE ONLY -- tests/golden/decoders.npz from the REAL reference decoders (given_models.py:159-189, 268-280).

    python -m oracle.make_golden_decoders      (build container: /root/reference mounted, torchaudio installed)

SpectrogramAE.decode (InverseSpectrogram) is deterministic and used as is.  MagSpectrogramAE / MelSpectrogramAE decode through
T.GriffinLim, whose default start is random phase: the fixtures replace the module's `decoder` by the same torchaudio transform
with rand_init=False and 8 iterations (the library's own code path with a reproducible start); the InverseMelScale output
(`inv_melscale_t`, the reference's own instance) is stored separately.
"""
import os
import sys

import numpy as np
import torch
import torchaudio.transforms as T

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.refload import load_reference  # noqa: E402
from oracle.make_golden import synth  # noqa: E402


def main():
    gm = load_reference()["given_models"]
    torch.set_grad_enabled(False)
    d = {}
    x = synth((2, 2, 8192), 21)
    d["x"] = x.numpy()
    for tag, (n_fft, hop) in {"1024_256": (1024, 256), "2048_512": (2048, 512)}.items():
        m = gm.SpectrogramAE(n_fft=n_fft, hop_length=hop)
        spec = m.encode(x)
        d[f"roundtrip_{tag}"] = m.decode(spec).numpy()
        g = torch.Generator().manual_seed(5)
        z = torch.complex(torch.randn(3, n_fft // 2 + 1, 9, generator=g), torch.randn(3, n_fft // 2 + 1, 9, generator=g))
        d[f"rand_spec_{tag}"] = z.numpy()
        d[f"rand_istft_{tag}"] = m.decoder(z).numpy()
        mm = gm.MagSpectrogramAE(n_fft=n_fft, hop_length=hop)
        p = mm.encode(x)
        mm.decoder = T.GriffinLim(n_fft=n_fft, hop_length=hop, rand_init=False, n_iter=8)
        d[f"gl8_{tag}"] = mm.decode(p).numpy()
        me = gm.MelSpectrogramAE(sample_rate=48000, n_fft=n_fft, hop_length=hop)
        mel = me.encode(x)
        d[f"invmel_{tag}"] = me.inv_melscale_t(mel).numpy()
        me.decoder = T.GriffinLim(n_fft=n_fft, hop_length=hop, rand_init=False, n_iter=8)
        d[f"mel_gl8_{tag}"] = me.decode(mel).numpy()
    md = gm.MagDPhaseSpectrogramAE(n_fft=1024, hop_length=256)          # "Exact decoder" (unbatched [c, N] input)
    reps = md.encode(x[0])
    d["mdp_reps"] = reps.numpy()
    d["mdp_decode"] = md.decode(reps).numpy()
    out = os.path.join(ROOT, "tests", "golden", "decoders.npz")
    np.savez_compressed(out, **d)
    print(out, {k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    main()
