"""Parity of the BENCHED configuration at FULL depth (c2: 26 executed SigLIP layers x 2 images, 32 Llama-3-8B layers, 8 mask
regions, depth ON, S = 259): the CUDA path against the oracle fixture tests/golden/c2_full_depth.npz (written by
`tools/oracle_full.py oracle`: the CPU oracle in fp32 AND in its bf16 mode - the reference's own arithmetic - on the same seeded
weights and request).  Stated rule at depth 32: per-step logits no further from the fp32 oracle than 1.25 x the bf16 oracle is
(max-abs and rms), greedy ids exact wherever the top-1/top-2 margin exceeds 4 x that rms noise (vs both oracles), stage tensors
within 5e-2 rms, CUDA-graph decode == eager decode.  Needs ~20 GB of host memory and about two minutes to regenerate the 8B seeded weights."""
import os

import pytest

pytestmark = pytest.mark.gpu


def test_c2_full_depth_matches_oracle_fixture(golden_dir):
    from tools.oracle_full import check_against_fixture

    path = os.path.join(golden_dir, "c2_full_depth.npz")
    report, fails = check_against_fixture(path)
    if report.get("weights_match") is False:
        pytest.skip("torch's CPU generator produced different seeded weights on this host; the fixture does not apply")
    print(report)
    if os.environ.get("SRGPT_FULL_DEPTH_REPORT"):  # the committed artefact profiles/r02_oracle_c2_full.json is written this way
        import json
        with open(os.environ["SRGPT_FULL_DEPTH_REPORT"], "w") as f:
            json.dump(report, f, indent=1)
    assert not fails, fails
    # `fails` already holds the id rule (ids must agree wherever the oracle's margin exceeds 4 x the rms noise).  Near-ties are free to
    # flip with any change of summation order: step 2 of the fixture has a margin of 0.11 logits under an rms noise of 0.2.
    # (the fixture's steps 1 and 2 have margins of 0.27 and 0.11 logits: any change of an fp32 summation order - e.g. the region projector
    # GEMMs moving from split-k partials to whole tiles - may move the sequence off the oracle's there; step 0, margin 1.76, is robust)
    assert report["steps_compared_vs_fp32"] >= 2 and report["steps_equal_to_bf16_oracle"] >= report["steps_id_must_agree_bf16"]
