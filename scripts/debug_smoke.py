import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from ccphylo_b200 import api, synth

n, length = 70, 5000 + 13
codes = synth.make_codes(n, length, seed=7, nrun=0.05)
codes[3, :] = 4
seqs, masks, inc = oracle.encode_samples(codes)
min_len = max(1, int(0.5 * length))
include = (inc >= min_len).astype(np.uint8)
print("excluded:", np.where(include == 0)[0], "min_len", min_len)
gmask = oracle.global_mask(codes, include)
Dgo, dngo, ginco = oracle.fsa_cmp_global(seqs, gmask, include, length, norm=1000)
for trial in range(3):
    for use_ctx in (False, True):
        ctx = api.Context() if use_ctx else None
        Dg, _, dng, ginc = api.fsa_cmp_thread_out(seqs, include, gmask.reshape(1, -1), length, pair=False, norm=1000, ctx=ctx)
        bad = np.where(Dg != Dgo)[0]
        print(trial, use_ctx, dng, dngo, ginc, ginco, "nbad", len(bad), "of", len(Dg))
        if len(bad):
            print(" first bad cells", bad[:10], Dg[bad[:5]], Dgo[bad[:5]])
            # which rows
            rows = np.floor((1 + np.sqrt(1 + 8 * bad)) / 2).astype(int)
            print(" rows", np.unique(rows)[:20])
        if ctx: ctx.close()
