"""Minimal stand-in for the `texttable` package (not installed in this image; no
network), covering what GPU-BLOB's calculateOffloadThreshold.py uses
(reference calculateOffloadThreshold.py:5,56-58): Texttable(), add_rows(rows),
draw().  Put this directory on PYTHONPATH to run the UNMODIFIED script:

    PYTHONPATH=gpu-blas-offload-benchmark_b200/harness/pyshims \
        python /path/to/calculateOffloadThreshold.py cpu.csv gpu.csv
"""


class Texttable:
    def __init__(self, max_width=0):
        self._rows = []
        self._header = None

    def header(self, row):
        self._header = [str(c) for c in row]

    def add_row(self, row):
        self._rows.append([str(c) for c in row])

    def add_rows(self, rows, header=True):
        rows = list(rows)
        if header and rows:
            self.header(rows[0])
            rows = rows[1:]
        for r in rows:
            self.add_row(r)

    def draw(self):
        table = ([self._header] if self._header else []) + self._rows
        if not table:
            return ""
        ncol = max(len(r) for r in table)
        table = [r + [""] * (ncol - len(r)) for r in table]
        widths = [max(len(r[c]) for r in table) for c in range(ncol)]
        sep = "+" + "+".join("-" * (w + 2) for w in widths) + "+"
        hsep = "+" + "+".join("=" * (w + 2) for w in widths) + "+"
        lines = [sep]
        for i, r in enumerate(table):
            lines.append("| " + " | ".join(c.ljust(w) for c, w in zip(r, widths)) + " |")
            lines.append(hsep if (i == 0 and self._header) else sep)
        return "\n".join(lines)
