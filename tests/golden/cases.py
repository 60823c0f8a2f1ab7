"""Golden cases: BASELINE.json configs 1-5 scaled down (+ a FASTA and a ragged case).

Shared by make_golden.py (writes the fixtures from the reference) and the tests
(regenerate the same input bytes and compare against the stored reference output).
"""
import numpy as np

SQ_FIRST, SQ_BEST, SQ_ALL = 0, 1, 2
SQ_FAIL, SQ_CONVERT, SQ_IGNORE = 0, 4, 8


def fixed_pattern(seed: int, n: int) -> str:
    rng = np.random.default_rng(seed)
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


ALL_OPTS = [m | x for m in (SQ_FIRST, SQ_BEST, SQ_ALL) for x in (SQ_FAIL, SQ_CONVERT, SQ_IGNORE)]

CASES = {
    # cfg1: seeq -c -d 2 GATCGGAAGAGC on 150-nt reads
    "cfg1_small": dict(pattern="GATCGGAAGAGC", tau=2, reads=3000, options=ALL_OPTS,
                       gen=dict(seed=1, line_len=150, plant="GATCGGAAGAGC", plant_per_1024=102, max_edits=2)),
    # cfg2: seeq -b -l -p -k -d 1 A[CG]TNNGATC, reads with N
    "cfg2_small": dict(pattern="A[CG]TNNGATC", tau=1, reads=3000, options=ALL_OPTS,
                       gen=dict(seed=2, line_len=150, n_per_1024=5)),
    # cfg3: seeq -a -f -d 4 <40-mer> on 10-kb reads
    "cfg3_small": dict(pattern=fixed_pattern(3, 40), tau=4, reads=40, options=[SQ_ALL, SQ_BEST, SQ_FIRST],
                       gen=dict(seed=3, line_len=10_000, plant=fixed_pattern(3, 40), plant_per_1024=1024, max_edits=4)),
    # cfg4: 100-nt pattern, -d 8, -x 1, 250-nt reads with non-DNA bytes
    "cfg4_small": dict(pattern=fixed_pattern(4, 100), tau=8, reads=1500,
                       options=[SQ_BEST | SQ_CONVERT, SQ_FIRST | SQ_CONVERT, SQ_ALL | SQ_IGNORE, SQ_ALL | SQ_FAIL],
                       gen=dict(seed=4, line_len=250, plant=fixed_pattern(4, 100), plant_per_1024=102, max_edits=8,
                                junk_per_1024=1)),
    # cfg5: seeq -e -d 2 over FASTQ-like 4-line records
    "cfg5_small": dict(pattern="GATCGGAAGAGC", tau=2, reads=1500, options=[SQ_FIRST, SQ_BEST, SQ_ALL],
                       gen=dict(seed=5, line_len=150, plant="GATCGGAAGAGC", plant_per_1024=307, max_edits=2, fastq=True)),
    # FASTA: '>' headers are not counted (seeq.c:367-377); ragged lines, empty lines, no final newline
    "fasta_ragged": dict(pattern="GAT[CT]NCA", tau=1, reads=0, options=ALL_OPTS, gen=None, special="fasta_ragged"),
}


def make_input(B, case) -> np.ndarray:
    """The input bytes of a case, as a numpy uint8 array (host generator only)."""
    if case.get("special") == "fasta_ragged":
        rng = np.random.default_rng(77)
        parts = []
        for k in range(400):
            if k % 7 == 0:
                parts.append(b">seq%d some description\n" % k)
            n = int(rng.integers(0, 120))
            s = "".join("ACGTNacgtRY-"[i] for i in rng.choice(12, n, p=[.22, .22, .22, .22, .03, .01, .01, .01, .01, .02, .02, .01]))
            parts.append(s.encode() + b"\n")
        parts.append(b"GATCNCATTTGATTNCA")       # last line without '\n'
        return np.frombuffer(b"".join(parts), dtype=np.uint8).copy()
    g = B.make_gen(**case["gen"])
    return B.gen_host(g, case["reads"])


# CLI text-level cases: flags given to the (re-linked) reference front-end seeq-main.c.
# Parity is taken with split = 0: both the reference CLI
# (oracle/_ref/seeq_ref) and ours are built with -ftrivial-auto-var-init=zero.
CLI_CASES = {
    "cli_cfg1_count": ("cfg1_small", ["-c", "-d", "2"]),
    "cli_cfg2_best_lpk": ("cfg2_small", ["-b", "-l", "-p", "-k", "-d", "1"]),
    "cli_cfg3_all_compact": ("cfg3_small", ["-a", "-f", "-d", "4"]),
    "cli_cfg4_best_x1": ("cfg4_small", ["-b", "-f", "-x", "1", "-d", "8"]),
    "cli_cfg4_count_x1": ("cfg4_small", ["-c", "-x", "1", "-d", "8"]),
    "cli_cfg5_endline": ("cfg5_small", ["-e", "-d", "2"]),
    "cli_cfg5_prefix": ("cfg5_small", ["-r", "-d", "2"]),
    "cli_cfg1_matchonly": ("cfg1_small", ["-m", "-d", "2"]),
    "cli_cfg1_invert_lines": ("cfg1_small", ["-i", "-l", "-d", "2"]),
    "cli_cfg2_default": ("cfg2_small", ["-d", "1"]),
    "cli_fasta_default": ("fasta_ragged", ["-d", "1", "-x", "2"]),
    "cli_fasta_compact_all": ("fasta_ragged", ["-a", "-f", "-d", "1", "-x", "1"]),
    "cli_fasta_invert": ("fasta_ragged", ["-i", "-d", "1"]),
}
