#!/usr/bin/env python
"""Golden vectors from the reference's notebook FIGURES.

The notebooks under /root/reference/examples/notebooks store their plots as SVG (Plots.jl / GR backend): every series
is a <polyline> (or a run of <circle>s for scatter plots) in pixel coordinates with 6 significant digits, every axis
has its grid lines and tick labels in the same file.  Calibrating pixels against the grid lines turns the plotted
curves back into the numbers the reference computed -- to about 1e-5 relative on a log axis -- which pins the
oracle for every lattice and for cases the reference never prints as text.

    python tests/golden/extract_notebook_plots.py            # rewrites tests/golden/notebook_figures.json

Reads /root/reference (build container only); the JSON it writes is the committed fixture the tests use.
"""
import json
import os
import re
import sys

REF = "/root/reference/examples/notebooks"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "notebook_figures.json")

QUADRATURES = ["D2Q4", "D2Q5", "D2Q9", "D2Q13", "D2Q17", "D2Q21", "D2Q37"]  # LatticeBoltzmann.Quadratures, quadratures.jl:30-38


def cell_svg(notebook, cell):
    nb = json.load(open(os.path.join(REF, notebook)))
    for o in nb["cells"][cell]["outputs"]:
        if "data" in o and "image/svg+xml" in o["data"]:
            v = o["data"]["image/svg+xml"]
            return "".join(v) if isinstance(v, list) else v
    raise KeyError((notebook, cell))


def _points(attr):
    pts = re.search(r'points="([^"]*)"', attr, re.S).group(1).split()
    return [tuple(float(v) for v in p.split(",")) for p in pts]


def parse_panels(svg):
    """Split a Plots.jl/GR SVG into panels: grid-line pixel positions, tick labels, series (polylines and circles)."""
    panels, cur = [], None
    pending_rect = False
    for m in re.finditer(r"<(polyline|clipPath|text|circle|rect)\b([^>]*)>([^<]*)", svg):
        tag, attr, body = m.group(1), m.group(2), m.group(3)
        if tag == "clipPath":
            pending_rect = re.search(r'id="(\w+)"', attr).group(1)
            continue
        if tag == "rect":
            if pending_rect:
                clip_id, pending_rect = pending_rect, False
                x, y, w, h = (float(re.search(rf'\b{k}="([\d.]+)"', attr).group(1)) for k in ("x", "y", "width", "height"))
                if x > 0 and y > 0:  # a subplot's clip rectangle (the first two are canvas-sized)
                    cur = dict(clip=clip_id, rect=(x, y, w, h), xgrid=[], ygrid=[], xticks=[], yticks=[], title=None, series=[], _texts=[])
                    panels.append(cur)
            continue
        if cur is None:
            continue
        if tag == "polyline":
            pts = _points(attr)
            opacity = float(re.search(r"stroke-opacity:([\d.]+)", attr).group(1))
            color = re.search(r"stroke:(#[0-9a-f]+)", attr).group(1)
            if opacity == 0.1 and len(pts) == 2:  # grid line
                (x0, y0), (x1, y1) = pts
                (cur["xgrid"] if x0 == x1 else cur["ygrid"]).append(x0 if x0 == x1 else y0)
            elif color == "#000000" and len(pts) == 2 and not cur["_texts"] and not cur["series"]:
                pass  # axes and tick marks
            elif f"#{cur['clip']})" in attr:  # (legend samples are drawn with the canvas clip)
                cur["series"].append(dict(kind="line", color=color, opacity=opacity, px=pts))
        elif tag == "circle":
            cx, cy = (float(re.search(rf'\b{k}="([\d.eE+-]+)"', attr).group(1)) for k in ("cx", "cy"))
            fill = re.search(r"fill:(#[0-9a-f]+)", attr)
            fill = fill.group(1) if fill else None
            r = float(re.search(r'\br="([\d.]+)"', attr).group(1))
            if f"#{cur['clip']})" in attr:
                cur["series"].append(dict(kind="circle", color=fill, r=r, px=[(cx, cy)]))
        elif tag == "text":
            fs = int(re.search(r"font-size:(\d+)", attr).group(1))
            x = float(re.search(r'\bx="([\d.]+)"', attr).group(1))
            y = float(re.search(r'\by="([\d.]+)"', attr).group(1))
            cur["_texts"].append((fs, x, y, body))
    for p in panels:
        _ticks(p)
    return panels


def _ticks(p):
    """Tick labels -> numbers.  Log axes: '10' (big font) followed by the exponent pieces in a smaller font."""
    x0, y0, w, h = p["rect"]
    labels = []  # (x, y, value, is_log)
    texts = p.pop("_texts")
    big = max((fs for fs, *_ in texts if fs <= 60), default=48)
    i = 0
    while i < len(texts):
        fs, x, y, s = texts[i]
        if fs > 60:  # titles / axis names
            if p["title"] is None and y < y0:
                p["title"] = s
            i += 1
            continue
        if fs == big:
            j = i + 1
            exp = ""
            while j < len(texts) and texts[j][0] < big:
                exp += texts[j][3]
                j += 1
            s = s.strip()
            if exp.strip() and s.endswith("10"):
                mant = s[:-2].rstrip("×").strip()
                e = float(exp.strip())
                labels.append((x, y, e if not mant else None, True) if not mant else (x, y, float(mant) * 10 ** e, False))
            else:
                try:
                    labels.append((x, y, float(s), False))
                except ValueError:
                    pass
            i = j
        else:
            i += 1
    xl = [l for l in labels if l[1] > y0 + h + 40]                 # a text line below the plot area
    yl = [l for l in labels if l[1] <= y0 + h + 40 and l[0] < x0]  # left of it (the lowest one may hang below the axis)
    p["xticks"] = [(g, l[2], l[3]) for g, l in zip(p["xgrid"], xl)] if len(xl) == len(p["xgrid"]) else None
    p["yticks"] = [(g, l[2], l[3]) for g, l in zip(p["ygrid"], yl)] if len(yl) == len(p["ygrid"]) else None


def _axis(ticks):
    """pixel -> value map from >= 2 (pixel, value, is_log) ticks (least squares; exact for GR's linear mapping)."""
    import numpy as np
    px = np.array([t[0] for t in ticks])
    v = np.array([t[1] for t in ticks], dtype=float)
    a, b = np.polyfit(px, v, 1)
    log = ticks[0][2]
    return (lambda q: 10 ** (a * q + b)) if log else (lambda q: a * q + b)


def calibrated_series(panel):
    fx, fy = _axis(panel["xticks"]), _axis(panel["yticks"])
    out = []
    for s in panel["series"]:
        out.append(dict(kind=s["kind"], color=s["color"], points=[(fx(x), fy(y)) for x, y in s["px"]]))
    return out


def _group_circles(series):
    """consecutive circles of one colour = one scatter series (Plots draws an outline circle + a fill circle per marker)"""
    groups = []
    for s in series:
        if s["kind"] != "circle":
            continue
        if groups and groups[-1]["color"] == s["color"]:
            groups[-1]["points"] += s["points"]
        else:
            groups.append(dict(color=s["color"], points=list(s["points"])))
    return groups


def shear_wave_convergence():
    """shear_wave.ipynb cell 12 (code: cells 9-12 + notebook_examples.jl:34-69): static DecayingShearFlow(nu, scale),
    nu = tau / (2 css), tau = 0.8, SRT, AnalyticalEquilibrium, t_end = 1, TrackHydrodynamicErrors(problem, false,
    n_steps, NoStoppingCriteria()); one curve per quadrature over N = 8 scale, scale = 1, 2, 4, 8."""
    panels = parse_panels(cell_svg("shear_wave.ipynb", 12))
    names = ["error_u", "error_p", "error_sxy", "error_sxx"]
    out = {}
    for name, p in zip(names, panels):
        lines = [s for s in calibrated_series(p) if s["kind"] == "line" and s["color"] != "#808080"]
        if not lines:
            continue
        assert len(lines) == len(QUADRATURES), (name, len(lines))
        out[name] = {q: dict(N=[round(x) for x, _ in s["points"]], value=[y for _, y in s["points"]]) for q, s in zip(QUADRATURES, lines)}
    return dict(source="examples/notebooks/shear_wave.ipynb cell 12 (SVG polylines)", tau=0.8, scales=[1, 2, 4, 8], errors=out)


def shear_wave_d2q9():
    """shear_wave.ipynb cell 10: the same study for D2Q9 alone, error_u / error_p / error_sigma_xy over N."""
    (p,) = parse_panels(cell_svg("shear_wave.ipynb", 10))
    lines = [s for s in calibrated_series(p) if s["kind"] == "line" and s["color"] != "#808080"]
    assert len(lines) == 3
    return dict(source="examples/notebooks/shear_wave.ipynb cell 10", tau=0.8,
                errors={k: dict(N=[round(x) for x, _ in s["points"]], value=[y for _, y in s["points"]])
                        for k, s in zip(["error_u", "error_p", "error_sxy"], lines)})


def tgv_convergence():
    """taylor_green_vortex.ipynb cell 5 (code: cells 3-5): TGV(q, tau, scale, 31 scale, 17 scale, sqrt(0.01) / scale) on D2Q9,
    AnalyticalEquilibrium, ProcessingMethod(problem, false, t_end), simulate(model, 1:t_end), t_end =
    round(Int, decay_time(problem)); scatter markers ordered tau = 3.0, 2.0, 1.0, 0.8, each at scale = 1, 2, 4."""
    panels = parse_panels(cell_svg("taylor_green_vortex.ipynb", 5))
    taus, scales = [3.0, 2.0, 1.0, 0.8], [1, 2, 4]
    errors = {}
    for name, p in zip(["error_u", "error_p", "error_sxy", "error_sxx"], panels):
        pts = [s["points"][0] for s in calibrated_series(p) if s["kind"] == "circle"]
        assert len(pts) == len(taus) * len(scales), (name, len(pts))
        assert [round(x) for x, _ in pts] == [31 * sc for _ in taus for sc in scales]
        errors[name] = [[pts[i * len(scales) + j][1] for j in range(len(scales))] for i in range(len(taus))]
    return dict(source="examples/notebooks/taylor_green_vortex.ipynb cell 5 (SVG scatter markers)", taus=taus, scales=scales,
                NX=[31 * sc for sc in scales], NY=[17 * sc for sc in scales], errors=errors)


TGV_INIT_STRATEGIES = ["ConstantDensity", "AnalyticalVelocityAndStress", "AnalyticalEquilibrium",
                       "AnalyticalEquilibriumAndOffEquilibrium", "IterativeInitializationMeiEtAl(0.8, 1e-10)",
                       "IterativeInitializationMeiEtAl(1.0, 1e-10)"]


def tgv_init_strategies():
    """taylor_green_vortex.ipynb cell 9 (code: cells 7-9): TGV(D2Q9(), 0.8, 2, 96, 72, 0.03), one run per initialisation
    strategy, ProcessingMethod(problem, true, t_end) -> one TrackHydrodynamicErrors row per next! call of
    simulate(model, 1:t_end), t_end = 840: 841 rows.  Kept: rows 1-20, then every 20th, and the last."""
    panels = parse_panels(cell_svg("taylor_green_vortex.ipynb", 9))
    keep = sorted(set(list(range(0, 20)) + list(range(19, 841, 20)) + [840]))
    errors = {}
    for name, p in zip(["error_u", "error_p", "error_sxx", "error_sxy"], panels):
        lines = [s for s in calibrated_series(p) if s["kind"] == "line"]
        assert len(lines) == len(TGV_INIT_STRATEGIES) and all(len(s["points"]) == 841 for s in lines)
        assert all(abs(x - (i + 1)) < 0.01 + 1e-3 * (i + 1) for s in lines for i, (x, _) in enumerate(s["points"]))
        errors[name] = [[s["points"][i][1] for i in keep] for s in lines]
    return dict(source="examples/notebooks/taylor_green_vortex.ipynb cell 9 (SVG polylines, 841 points each)",
                strategies=TGV_INIT_STRATEGIES, rows=[i + 1 for i in keep], n_rows=841, errors=errors)


def couette_convergence():
    """couette.ipynb cell 7 (code: cell 6): CouetteFlow(1.0, u_0 / scale, nu, 1, 5 scale, (1.0, 1.0)), nu = tau / (2 css),
    tau = 0.8, SRT, ZeroVelocityInitialCondition, t_end = 1, TrackHydrodynamicErrors(problem, false, n_steps,
    VelocityConvergenceStoppingCriteria(1e-7, problem)); error_u per quadrature over scale = 1, 2, 4, 8 (plotted at
    x = 8 scale), one panel per u_0.  (The stored figure has 4 scales; the code cell was later edited to 6.)"""
    panels = parse_panels(cell_svg("couette.ipynb", 7))
    us = [0.01, 0.015, 0.02, 0.03, 0.06, 0.12]
    lattices = ["D2Q9", "D2Q13", "D2Q17", "D2Q21", "D2Q37"]
    assert len(panels) == len(us)
    out = {}
    for u0, p in zip(us, panels):
        lines = [s for s in calibrated_series(p) if s["kind"] == "line" and s["color"] != "#808080"]
        assert len(lines) == len(lattices)
        assert all([round(x) for x, _ in s["points"]] == [8, 16, 32, 64] for s in lines)
        out[str(u0)] = {l: [y for _, y in s["points"]] for l, s in zip(lattices, lines)}
    return dict(source="examples/notebooks/couette.ipynb cell 7 (SVG polylines)", tau=0.8, u_0=us, lattices=lattices,
                scales=[1, 2, 4, 8], error_u=out)


def poiseuille_tau_sweep():
    """poiseuille.ipynb cell 9 (code: cells 6-9): error_u of the D2Q9 TRT(tau, tau, force) Poiseuille solve (the diagonal of
    the trt_magic_parameter study, PoiseuilleFlow((tau - 0.5) / 3, 1), ZeroVelocityInitialCondition, t_end = 100,
    VelocityConvergenceStoppingCriteria(1e-7)) for tau = 0.51, 0.52, ..., 10.0."""
    (p,) = parse_panels(cell_svg("poiseuille.ipynb", 9))
    (s,) = calibrated_series(p)
    assert len(s["points"]) == 950 and all(abs(x - (0.51 + 0.01 * i)) < 2e-4 for i, (x, _) in enumerate(s["points"]))
    return dict(source="examples/notebooks/poiseuille.ipynb cell 9 (SVG polyline, 950 points)",
                tau=[round(0.51 + 0.01 * i, 2) for i in range(950)], error_u=[y for _, y in s["points"]])


def _profile_circles(panel, n_snapshots, nx):
    pts = [s["points"][0] for s in calibrated_series(panel) if s["kind"] == "circle"]
    assert len(pts) == n_snapshots * nx, len(pts)
    return [[y for _, y in pts[k * nx:(k + 1) * nx]] for k in range(n_snapshots)], [x for x, _ in pts[:nx]]


def shear_wave_snapshots():
    """shear_wave.ipynb cells 3-4 and 6-7 (plot_snapshots, notebook_examples.jl:71-240): dimensionless sigma_xx and sigma_xy
    along x at y_pos = round(Int, NY / 2) for the TakeSnapshots snapshots of
      decaying: DecayingShearFlow(1/6, 4, A = 3.0, static = false), AnalyticalEquilibrium, snapshots at steps
                round.(Int, [0, 0.05, 0.15, 0.25] ./ dt) .+ 1   (a travelling, decaying wave; no force)
      static:   DecayingShearFlow(1/6, 16, A = 0.5, static = true), ZeroVelocityInitialCondition, snapshots at steps
                round.(Int, [0.01, 0.1, 1, 10] ./ (nu dt))       (spin-up under the TIME-DEPENDENT force)."""
    out = {}
    for key, cell, nx in (("decaying", 4, 32), ("static", 7, 128)):
        panels = parse_panels(cell_svg("shear_wave.ipynb", cell))
        sxx, x = _profile_circles(panels[2], 4, nx)
        sxy, _ = _profile_circles(panels[3], 4, nx)
        out[key] = dict(NX=nx, x=x, sigma_xx=sxx, sigma_xy=sxy)
    return dict(source="examples/notebooks/shear_wave.ipynb cells 4 and 7 (SVG scatter markers)", **out)


def wall_snapshots():
    """poiseuille.ipynb cells 3-4 and couette.ipynb cells 2-4 (plot_snapshots methods for PoiseuilleFlow / CouetteFlow,
    notebook_examples.jl:227-540): dimensionless sigma_xx and sigma_xy along y at x_pos = max(round(Int, NX / 2), 1) of the
    TakeSnapshots snapshots of the spin-up from rest, D2Q9, SRT:
      poiseuille: PoiseuilleFlow(1/6, 4) (3 x 20), snapshots at steps round.(Int, [0.01, 0.05, 0.1, 1.0] ./ (nu dt))
      couette:    CouetteFlow(1/6, 16) (1 x 80), snapshots at round.(Int, [0, 0.005, 0.01, 0.05, 0.1, 0.5, 1, 5] ./ (nu dt)) .+ 1"""
    out = {}
    for key, nbk, n_snap, ny in (("poiseuille", "poiseuille.ipynb", 4, 20), ("couette", "couette.ipynb", 8, 80)):
        panels = parse_panels(cell_svg(nbk, 4))
        sxx, y = _profile_circles(panels[2], n_snap, ny)
        sxy, _ = _profile_circles(panels[3], n_snap, ny)
        out[key] = dict(NY=ny, y=y, sigma_xx=sxx, sigma_xy=sxy)
    return dict(source="examples/notebooks/poiseuille.ipynb cell 4, couette.ipynb cell 4 (SVG scatter markers)", **out)


def main():
    if not os.path.isdir(REF):
        sys.exit("the reference notebooks are not available here; the committed JSON is the fixture")
    fixtures = dict(
        _comment="extracted by tests/golden/extract_notebook_plots.py from the SVG figures of the reference's notebooks; "
                 "values carry ~1e-5 relative calibration error",
        shear_wave_convergence=shear_wave_convergence(),
        shear_wave_d2q9=shear_wave_d2q9(),
        tgv_convergence=tgv_convergence(),
        tgv_init_strategies=tgv_init_strategies(),
        couette_convergence=couette_convergence(),
        poiseuille_tau_sweep=poiseuille_tau_sweep(),
        shear_wave_snapshots=shear_wave_snapshots(),
        wall_snapshots=wall_snapshots(),
    )
    fixtures = json.loads(json.dumps(fixtures), parse_float=lambda v: float("%.7g" % float(v)))  # 7 digits are plenty
    json.dump(fixtures, open(OUT, "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
