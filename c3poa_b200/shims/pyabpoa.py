"""Drop-in for the part of pyabpoa 1.0.5 the reference uses (bin/determine_consensus.py:30-47):
    res = msa_aligner(match=5).msa(seqs, out_cons=True, out_msa=True); res.cons_seq[0]
The alignment and consensus run on the GPU (c3_poa_batch).  MSA rows are produced for exactly two
sequences with out_cons=False (the only case in which the reference reads msa_seq: its 2-repeat path)."""
from ..api import GpuConsensus, default_poa_params

_GPU = None


class msa_result:
    def __init__(self, n_seq, cons, msa=()):
        self.n_seq = n_seq
        self.n_cons = 1 if cons else 0
        self.cons_len = [len(cons)] if cons else []
        self.cons_seq = [cons] if cons else []
        self.msa_seq = list(msa)
        self.msa_len = len(self.msa_seq[0]) if self.msa_seq else 0


class msa_aligner:
    def __init__(self, aln_mode="g", is_aa=False, match=2, mismatch=4, score_matrix=b"", gap_open1=4, gap_open2=24,
                 gap_ext1=2, gap_ext2=1, extra_b=10, extra_f=0.01, is_diploid=False, min_freq=0.3):
        if aln_mode != "g" or is_aa or score_matrix or is_diploid:
            raise NotImplementedError("only the global nucleotide mode the reference uses is implemented")
        self.params = default_poa_params(match=match, mismatch=mismatch, gap_open1=gap_open1, gap_ext1=gap_ext1,
                                         gap_open2=gap_open2, gap_ext2=gap_ext2, wb=extra_b, wf=extra_f)

    def msa(self, seqs, out_cons, out_msa, out_pog=b"", incr_fn=b""):
        global _GPU
        if _GPU is None:
            _GPU = GpuConsensus(0)
        if not seqs:
            return msa_result(0, "")
        pair = out_msa and not out_cons and len(seqs) == 2
        r = _GPU.poa_batch([list(seqs)], params=self.params, want_msa=pair)
        if r["status"][0] != 0:
            raise RuntimeError(f"GPU POA failed with status {int(r['status'][0])}")
        return msa_result(len(seqs), r["cons"][0] if out_cons else "", r["msa"][0] if pair else ())
