"""The C host program (5g-nr-randomaccess_b200/host/rach_sim.c) against the reference's own
stdout report and result files (RandomAccessWithNOMA.c:354-361, 741-825) produced by the
tape-mode reference build -- byte for byte."""
import importlib
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REF_SCRIPT = r'''
import sys
sys.path.insert(0, %r)
from oracle import oracle as O
cfg = O.make_config(nUE=int(sys.argv[1]), rep=int(sys.argv[2]), echo=2, maxMsg2TxCount=int(sys.argv[3]) - 1,
                    nGrantUL=int(sys.argv[4]), distribution=int(sys.argv[5]))
O.run_ref("w", cfg, per_ue=False)
''' % ROOT


def _reference(tmp, nue, rep, retx, grant, dist):
    d = tmp / ("ref_%d_%d" % (nue, rep))
    (d / "NomaBetaResults").mkdir(parents=True)
    (d / "NomaUniformResults").mkdir(parents=True)
    out = subprocess.run([sys.executable, "-c", REF_SCRIPT, str(nue), str(rep), str(retx), str(grant), str(dist)],
                         cwd=d, capture_output=True, text=True, check=True).stdout
    return d, out


@pytest.mark.parametrize("dist,sub", [(0, "NomaBetaResults"), (1, "NomaUniformResults")])
def test_stdout_and_files_match_reference(tmp_path, oracle, dist, sub):
    if not oracle.ref_available("w"):
        pytest.skip("oracle/_ref not shipped")
    pkg = importlib.import_module("5g-nr-randomaccess_b200")
    exe = pkg.build_host()
    nues, retx, grant = [1500, 4000], 5, 7
    out = subprocess.run([exe, "-t", "2", "-d", str(dist), "-mrc", str(retx), "-g", str(grant), "--nue",
                          ",".join(map(str, nues)), "--outdir", str(tmp_path / "ours")],
                         capture_output=True, text=True, check=True).stdout
    blocks = out.split("-------- ")
    assert blocks[0] == ("Traffic model: Uniform\n\n" if dist == 1 else "Traffic model: Beta\n\n")
    k = 1
    for seed in (0, 1):
        for nue in nues:
            d, ref_out = _reference(tmp_path, nue, seed, retx, grant, dist if dist == 1 else 0)
            ref_block = ref_out.split("-------- ")[1]
            assert blocks[k] == ref_block, (seed, nue)
            k += 1
            for kind in ("%d_Results.txt" % nue, "UE%05d_Logs.txt" % nue):
                ours = (tmp_path / "ours" / sub / ("%d_54_%s" % (seed, kind))).read_bytes()
                ref = (d / sub / ("0_54_%s" % kind)).read_bytes()
                assert ours == ref, (seed, nue, kind)


def test_cli_errors_match_reference_behaviour():
    pkg = importlib.import_module("5g-nr-randomaccess_b200")
    exe = pkg.build_host()
    r = subprocess.run([exe, "-p", "0"], capture_output=True, text=True)
    assert r.returncode == 255 and r.stdout == "Number of preamble must be greater than zero."     # W:106-108
    r = subprocess.run([exe, "-d", "2"], capture_output=True, text=True)
    assert r.returncode == 255 and r.stdout == "Traffic model just choose 1 or 2"                  # W:100-102
    r = subprocess.run([exe, "--bogus", "1"], capture_output=True, text=True)
    assert r.returncode == 255 and r.stdout.startswith("--times         -t : Simulation times (int)\n")
    assert "--hut           -u : Height of UE from ground (float)\n" in r.stdout                    # W:200
