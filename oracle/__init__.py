"""CPU oracle for the SVT_SpeechBrain AMT inference hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE.  It is a plain fp32 CPU restatement of
the reference algorithm (torch CPU ops / numpy / C), used only as the checker:
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.  Nothing under `svt_speechbrain_b200/` may
import, call or link anything from here; the product path has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against the reference's OWN modules imported from
/root/reference (`oracle/ref_bootstrap.py`, only possible in the authoring
container) -- see `oracle/make_golden.py`, which wrote `tests/golden/*.npz`.
The AV-HuBERT transformer body (fairseq, un-vendored, un-pinned) cannot be
imported anywhere: that part of the oracle is "parity unpinned".
"""
