"""Which switch breaks batch-16 parity?  generate(mel16) vs the oracle's robust cases under engine option combinations."""
import itertools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.gpu_common import PROMPT, mel_inputs, model_pair, robust_cases  # noqa: E402

dims, oracle, h = model_pair()
mel = mel_inputs(16)
P = np.asarray([PROMPT] * 16, np.int32)
for beam in (1, 5):
    res, robust = robust_cases(oracle, mel, [PROMPT] * 16, beam, n_probe=2)
    want = [r.sequences_ids[0] for r in res]
    print("beam", beam, "robust", robust, "lens", [len(w) for w in want], flush=True)
    for tc, pdl, graphs in itertools.product((1, 0), (1, 0), (1, 0)):
        h.set_option("cross_tc", tc)
        h.set_option("batch_pdl", pdl)
        h.set_option("use_graphs", graphs)
        got, _ = h.generate(mel, P, beam_size=beam)
        got2, _ = h.generate(mel, P, beam_size=beam)
        bad = [(i, next((k for k, (a, b) in enumerate(zip(got[i], want[i])) if a != b), -1)) for i in robust if got[i] != want[i]]
        print(f"  cross_tc={tc} pdl={pdl} graphs={graphs}: mismatches (utt, first index) {bad}; repeat run identical: {got == got2}", flush=True)
h.set_option("cross_tc", 1); h.set_option("batch_pdl", 1); h.set_option("use_graphs", 1)
