"""GPU parity across the marx.par parameter surface: for every parameter variation the reference's own initialisation
(oracle/_ref/calpack_dump = the stock *_init functions) produces the tables, the CPU restatement (pinned bit-exact to
the reference for these very cases by tests/test_oracle_vs_reference.py) traces 2^17 rays, and the CUDA path must agree
slot by slot at every stage boundary.  Needs oracle/_ref (it travels to the GPU box with the snapshot)."""
import os
import subprocess

import numpy as np
import pytest

from tests.test_gpu_oracle import check_compacted_equals_in_place, check_cuda_against_oracle
from tests.test_oracle_vs_reference import CASES, HAVE_REF, REF, expand_args

pytestmark = pytest.mark.gpu

EXTRA = [
    ("det_ideal", ["MinEnergy=0.5", "MaxEnergy=6.0", "GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=INTERNAL", "DetIdeal=yes"], 31, 0),
    ("no_blur_vignetting", ["MinEnergy=0.5", "MaxEnergy=6.0", "GratingType=NONE", "DetectorType=ACIS-S", "DitherModel=INTERNAL",
                            "HRMA_Use_Blur=no", "HRMAVig=0.8"], 32, 0),
    ("no_scale_factors_big_aspect_blur", ["MinEnergy=1.0", "MaxEnergy=7.0", "GratingType=HETG", "DetectorType=ACIS-S", "DitherModel=INTERNAL",
                                          "HRMA_Use_Scale_Factors=no", "AspectBlur=1.5", "P1Blur=0.6", "H6Blur=0.2"], 33, 65536),
    ("roll_dither_off_axis", ["MinEnergy=0.5", "MaxEnergy=4.0", "GratingType=NONE", "DetectorType=ACIS-S", "DitherModel=INTERNAL",
                              "DitherAmp_Roll=30", "DitherPeriod_Roll=700", "SourceRA=250.05", "Roll_Nom=17.0"], 34, 0),
    ("line_source_hetg", ["SourceType=LINE", "S-LinePhi=60", "S-LineTheta=120", "MinEnergy=0.8", "MaxEnergy=3.0", "GratingType=HETG",
                          "DetectorType=ACIS-S", "DitherModel=INTERNAL"], 35, 0),
    ("image_source_acis_s", ["SourceType=IMAGE", "S-ImageFile=%IMAGE%", "MinEnergy=0.5", "MaxEnergy=3.0", "GratingType=NONE",
                             "DetectorType=ACIS-S", "DitherModel=NONE"], 36, 0),
    ("hrc_i_letg", ["MinEnergy=0.1", "MaxEnergy=1.5", "GratingType=LETG", "DetectorType=HRC-I", "DitherModel=INTERNAL"], 37, 0),
    ("detector_none", ["MinEnergy=0.5", "MaxEnergy=4.0", "GratingType=HETG", "DetectorType=NONE", "DitherModel=NONE"], 38, 0),
]
_seen = {c[0] for c in CASES}
ALL = [(c[0], c[1], c[2], c[3]) for c in CASES] + [e for e in EXTRA if e[0] not in _seen and e[0] != "no_scale_factors_big_aspect_blur"]


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (compiled reference) not present on this box")
@pytest.mark.parametrize("name,args,seed,first", ALL, ids=[c[0] for c in ALL])
def test_cuda_matches_oracle_for_parameter_variation(tmp_path, name, args, seed, first):
    par = "@@" + os.path.join(REF, "par", "marx.par")
    common = ["ExposureTime=0", "Verbose=0", "SourceFlux=0.003", "TStart=2023.5", "SpectrumType=FLAT"]
    env = dict(os.environ, MARX_DATA_DIR=os.path.join(REF, "data"))
    # 2^17 rays at the default flux take ~38 ks: stretch the synthetic aspect solution (the ends_early case stays short)
    args = expand_args(tmp_path, [a.replace("%ASPSOL:6000%", "%ASPSOL:45000%") for a in args])
    pack = str(tmp_path / (name + ".calpack"))
    subprocess.check_call([os.path.join(REF, "calpack_dump"), pack, par] + common + args, env=env, stdout=subprocess.DEVNULL)
    counts = check_cuda_against_oracle(pack, seed, first, 1 << 17)
    print(name, counts)
    if "ends_early" in name:
        assert 0 < counts[0] < 1 << 17 and counts[1] > 0
    else:
        assert counts[0] == 1 << 17 and counts[1] > 0
    # ... and the compacting product path (other kernel cuts, fused entry) leaves exactly those survivors
    check_compacted_equals_in_place(pack, seed, first, 1 << 17)
