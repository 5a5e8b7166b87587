"""Seeded inputs shared by make_golden.py (runs the reference, here) and the tests (everywhere).
Every input is a pure function of integers via public_kssd_b200.synth (splitmix64), so the bytes are
identical on every box; the golden .npz files hold what the UNMODIFIED reference produced for them."""
import numpy as np

from public_kssd_b200 import synth

SHUF_SEED_S6 = 1      # subk 6 table used by L3K10 / L3K11 cases
SHUF_SEED_S5 = 2      # subk 5 table used by L2K8 cases
SHUF_ID = 4242


def fasta_inputs():
    """name -> uint8 FASTA text.  Names sort in a fixed order; the reference shuffles its own order
    (command_shuffle.c:135), so results are always keyed by file name."""
    g = {}
    g["a_plain80"] = synth.to_fasta(synth.random_bases(300_000, 11), "plain", 80)
    g["b_oneline"] = synth.to_fasta(synth.random_bases(200_000, 12), "one", 0)
    g["c_crlf70"] = synth.to_fasta(synth.random_bases(150_000, 13), "crlf", 70, crlf=True)
    g["d_messy"] = synth.messy_fasta(400_000, 3)
    g["e_messy_crlf"] = synth.messy_fasta(300_000, 4, crlf=True, width=70)
    g["f_short_lines"] = synth.to_fasta(synth.random_bases(60_000, 14), "w7", 7)
    b = synth.random_bases(200_000, 21)
    g["g_dup"] = synth.to_fasta(np.concatenate([b, b[50_000:150_000], synth.random_bases(1000, 22)]), "dup", 80)
    anc = synth.random_bases(250_000, 31)
    g["h_anc"] = synth.to_fasta(anc, "anc", 80)
    g["i_mut1"] = synth.to_fasta(synth.mutate(anc, 0.01, 32), "mut1", 80)
    g["j_mut5"] = synth.to_fasta(synth.mutate(anc, 0.05, 33), "mut5", 80)
    return g


def fastq_inputs():
    src = synth.random_bases(100_000, 41)
    g = {}
    g["a_cov20"] = synth.to_fastq(src, 13000, 150, seed=42)
    g["b_cov5_nonl"] = synth.to_fastq(src, 3300, 150, seed=43, trailing_newline=False)
    g["c_short"] = synth.to_fastq(src, 2000, 36, seed=44)
    return g


def byread_inputs():
    """FASTA-formatted read files for `--byread` (reference reads2mco, iseq2comem.c:78-186)."""
    src = synth.random_bases(400_000, 51)
    g = {}
    g["a_clean"] = synth.to_read_fasta(src, 900, seed=52)
    g["b_messy"] = synth.to_read_fasta(src, 700, seed=53, messy=True)
    g["c_leading_short"] = synth.to_read_fasta(src, 400, seed=54, min_len=5, max_len=60, width=0, leading_sequence=True)
    g["d_contigs"] = synth.messy_fasta(200_000, 6)
    return g
