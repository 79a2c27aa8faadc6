"""tests/golden/make_golden.py -- regenerates the committed golden fixtures.  Runs ONLY in the
build container (needs /root/reference and oracle/_ref built from it); the fixtures it writes are
what travels to the GPU box.

  pcg32_kat.json          the reference's own known-answer vectors for the one engine it uses:
                          pcg-cpp/test-high/expected/check-pcg32.out and check-pcg32_oneseq.out
                          (CRLF stripped), plus jump-ahead vectors drawn from the vendored header
                          through oracle/_ref.
  default_tables.npz      setup tables of the running reference object on its default plane.
  small_step.npz          three injected steps (constructor step + 2 filter steps) of the reference
                          object reshaped to a small synthetic plane: inputs and every output.
  default_step_digest.json  sha256 + probes of the reference's outputs for one injected step on the
                          default 510x400 plane (noise = oracle counter-based stream, seed 20261018).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import _dfb_import  # noqa: E402,F401
from digital_filtering_b200 import workloads as W  # noqa: E402
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/digital-filtering-c++/pcg-cpp/test-high/expected"


def kat():
    out = {}
    for name in ("check-pcg32", "check-pcg32_oneseq"):
        txt = open(os.path.join(REF, name + ".out"), "rb").read().decode().replace("\r", "")
        out[name] = txt.split("\n\n", 1)[1]          # drop the 5-line banner (typedef-specific size/period)
    jumps = []
    for delta in (0, 1, 6, 1000, 2 ** 32, 0xDEADBEEFCAFE, 2 ** 64 - 6):
        jumps.append(dict(delta=str(delta),
                          two_arg=[int(x) for x in O.ref_pcg32_draw(42, 54, delta, 4, True)],
                          one_arg=[int(x) for x in O.ref_pcg32_draw(42, 0, delta, 4, False)]))
    out["jump_ahead"] = jumps
    json.dump(out, open(os.path.join(HERE, "pcg32_kat.json"), "w"), indent=1)


def default_tables():
    R = O.RefFilter()
    P = R.plane()
    np.savez_compressed(os.path.join(HERE, "default_tables.npz"), Ny=P["Ny"], Nz=P["Nz"], rows=P["rows"], scales=P["scales"],
                        yc=P["yc"][:, 0], dy=P["dy"][:, 0], N_y=P["N_y"][:, :, 0], N_z=P["N_z"][:, :, 0],
                        Ny_max=P["Ny_max"], Nz_max=P["Nz_max"], u_tau=R.scalar(2), tau_w=R.scalar(3))
    # one injected step on the full default plane -> digest
    seed = 20261018
    Ny, Nz = P["Ny"], P["Nz"]
    rys = [O.noise_rys(seed, 0, f, 1, Ny, P["Ny_max"][f], Nz) for f in range(3)]
    hal = [O.noise_halo(seed, 0, f, 1, Ny, P["Nz_max"][f]) for f in range(3)]
    fo = [O.noise_elements(seed, 100 + f, 0, Ny * Nz, 0, Ny * Nz).reshape(Ny, Nz) for f in range(3)]
    R.inject(rys, hal, fo)
    R.step_injected(1e-5)
    o = R.outputs()
    dig = {}
    for k in ("filt", "fluc", "T", "rho"):
        a = np.ascontiguousarray(o[k])
        dig[k] = dict(sha256=hashlib.sha256(a.tobytes()).hexdigest(), rms=float(np.sqrt((a ** 2).mean())),
                      probe=[float(x) for x in a.ravel()[[0, 12345, 99999, a.size - 1]]])
    json.dump(dict(seed=seed, dt=1e-5, outputs=dig), open(os.path.join(HERE, "default_step_digest.json"), "w"), indent=1)
    R.close()


def small_step():
    plane = W.plane_profile(40, 36, 8, 6)
    R = O.RefFilter()
    R.reshape(plane)
    P = R.plane()
    seed = 7
    Ny, Nz = P["Ny"], P["Nz"]
    rec = dict(Ny=Ny, Nz=Nz, d_i=plane["d_i"], yc=plane["yc"], dy=plane["dy"], dz=plane["dz"], rows=plane["rows"],
               scales=plane["scales"], N_y=P["N_y"][:, :, 0], N_z=P["N_z"][:, :, 0], seed=seed, dts=np.array([0.0, 2e-7, 5e-7]))
    R.set_fvec(0, "filt_old", np.zeros(Ny * Nz)); R.set_fvec(1, "filt_old", np.zeros(Ny * Nz)); R.set_fvec(2, "filt_old", np.zeros(Ny * Nz))
    for s, dt in enumerate(rec["dts"]):
        rys = [O.noise_rys(seed, 0, f, s, Ny, P["Ny_max"][f], Nz) for f in range(3)]
        hal = [O.noise_halo(seed, 0, f, s, Ny, P["Nz_max"][f]) for f in range(3)]
        R.inject(rys, hal)
        if s == 0:
            R.first_step_injected()
        else:
            R.step_injected(float(dt))
        o = R.outputs()
        for f in range(3):
            rec[f"s{s}_rys{f}"] = rys[f]
            rec[f"s{s}_halo{f}"] = hal[f]
        for k in ("filt", "fluc", "T", "rho"):
            rec[f"s{s}_{k}"] = o[k]
    np.savez_compressed(os.path.join(HERE, "small_step.npz"), **rec)
    R.close()


if __name__ == "__main__":
    O.build()
    kat()
    default_tables()
    small_step()
    print("golden fixtures written to", HERE)
