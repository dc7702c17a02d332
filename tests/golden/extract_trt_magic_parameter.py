"""Extracts the only numeric golden table the reference ships -- the printed
output of examples/notebooks/trt_magic_parameter.ipynb (cell 3: 902,500-row
DataFrame of D2Q9 TRT+force Poiseuille errors; HTML view shows rows 1-30, the
text view rows 1-10 and the last 11) -- into trt_magic_parameter.json.

Run in the build container only (needs /root/reference):
    python tests/golden/extract_trt_magic_parameter.py
"""
import json
import re
import sys

NB = "/root/reference/examples/notebooks/trt_magic_parameter.ipynb"


def main():
    nb = json.load(open(NB))
    rows = {}
    for cell in nb["cells"]:
        for out in cell.get("outputs", []):
            data = out.get("data", {})
            html = "".join(data.get("text/html", []))
            for m in re.finditer(
                    r"<tr><th>(\d+)</th><td>([^<]+)</td><td>([^<]+)</td><td>\"D2Q9\"</td>"
                    r"<td>([^<]+)</td><td>([^<]+)</td><td>([^<]+)</td><td>([^<]+)</td></tr>", html):
                i, ts, ta, eu, ep, sxx, sxy = m.groups()
                rows[int(i)] = dict(row=int(i), tau_s=float(ts), tau_a=float(ta), error_u=float(eu),
                                    error_p=float(ep), error_sxx=float(sxx.replace("Inf", "inf")),
                                    error_sxy=float(sxy))
            text = "".join(data.get("text/plain", []))
            for m in re.finditer(r"│ (\d+)\s+│ ([\d.]+)\s+│ ([\d.]+)\s+│ \"D2Q9\"\s+│ ([\d.e-]+)\s+│ ([\d.e-]+)\s+│", text):
                i, ts, ta, eu, ep = m.groups()
                rows.setdefault(int(i), dict(row=int(i), tau_s=float(ts), tau_a=float(ta),
                                             error_u=float(eu), error_p=float(ep)))
    out = dict(
        source="examples/notebooks/trt_magic_parameter.ipynb:109-176 (cell 3 output)",
        setup=("D2Q9; PoiseuilleFlow(nu=(tau_s-0.5)/3, scale=1) -> NX=3, NY=5, u_max=0.1; "
               "TRT(tau_s, tau_a, lattice_force closure); ZeroVelocityInitialCondition; "
               "TrackHydrodynamicErrors(problem, false, 5000, VelocityConvergenceStoppingCriteria(1e-7)); "
               "t_end=100; value = processing_method.df[end]"),
        printed_digits=6,
        rows=[rows[k] for k in sorted(rows)],
    )
    json.dump(out, open(__file__.replace("extract_trt_magic_parameter.py", "trt_magic_parameter.json"), "w"), indent=1)
    print(len(out["rows"]), "rows")


if __name__ == "__main__":
    sys.exit(main())
