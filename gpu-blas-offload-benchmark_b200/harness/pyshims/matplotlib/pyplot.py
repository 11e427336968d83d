"""pyplot stand-in: one figure, one axes, categorical x (the reference plots strings),
line series with markers, horizontal guide lines, text, ticks, legend, PNG output."""
import math

import numpy as np
from PIL import Image, ImageDraw, ImageFont

_fig = None
_NAMED = {"black": (0, 0, 0), "white": (255, 255, 255)}


def _rgb(c):
    if isinstance(c, str) and c.startswith("#"):
        return tuple(int(c[i:i + 2], 16) for i in (1, 3, 5))
    return _NAMED.get(c, (0, 0, 0))


def _font(size):
    try:
        return ImageFont.load_default(size=max(8, int(size * 1.6)))
    except TypeError:  # old Pillow: fixed-size bitmap font
        return ImageFont.load_default()


class _Handle:
    def set_linewidth(self, w): pass
    def set_markersize(self, s): pass
    def set_markeredgewidth(self, w): pass


class _Legend:
    def __init__(self, n):
        self.legendHandles = [_Handle() for _ in range(n)]
        self.legend_handles = self.legendHandles


class _Axes:
    def __init__(self):
        self.series, self.hlines, self.texts = [], [], []

    def plot(self, x, y, color="black", marker=None, label=None, **kw):
        self.series.append((list(range(len(y))) if x and isinstance(x[0], str) else [float(v) for v in x],
                            [float(v) for v in y], _rgb(color), marker, label))

    def axhline(self, y, color="black", linestyle="-", **kw):
        self.hlines.append((float(y), _rgb(color), linestyle))

    def text(self, x, y, s, fontsize=12, **kw):
        self.texts.append((float(x), float(y), str(s), fontsize))


class _Figure:
    def __init__(self, figsize):
        self.figsize = figsize
        self.ax = _Axes()
        self.xticks = ([], [], 12)
        self.yticks = (None, 12)
        self.ylim = None
        self.xlabel = self.ylabel = self.title = ""
        self.label_fs = 12
        self.legend = None

    def add_subplot(self, *a, **kw):
        return self.ax

    def data_ymax(self):
        vals = [v for s in self.ax.series for v in s[1]] + [h[0] for h in self.ax.hlines]
        return max(vals) if vals else 1.0

    def auto_yticks(self):
        top = max(self.data_ymax(), 1e-12)
        raw = top / 8.0
        mag = 10 ** math.floor(math.log10(raw))
        step = min((m for m in (1, 2, 2.5, 5, 10) if m * mag >= raw), default=10) * mag
        n = int(math.ceil(top / step)) + 1
        return np.array([step * i for i in range(0, n + 1)])


def figure(figsize=(8, 6), **kw):
    global _fig
    _fig = _Figure(figsize)
    return _fig


def _cur():
    return _fig if _fig is not None else figure()


def xticks(ticks=None, labels=None, fontsize=12, **kw):
    f = _cur()
    if ticks is not None:
        f.xticks = (list(ticks), [str(l) for l in (labels if labels is not None else ticks)], fontsize)
    return np.array(f.xticks[0]), f.xticks[1]


def yticks(ticks=None, fontsize=12, **kw):
    f = _cur()
    if ticks is None:
        locs = f.auto_yticks()
        return locs, [str(v) for v in locs]
    f.yticks = (list(ticks), fontsize)
    return np.array(f.yticks[0]), [str(v) for v in f.yticks[0]]


def ylim(lo=None, hi=None):
    _cur().ylim = (lo, hi)


def margins(**kw): pass


def legend(fontsize=12, **kw):
    f = _cur()
    f.legend = fontsize
    return _Legend(len(f.ax.series))


def xlabel(s, fontsize=12, **kw):
    f = _cur(); f.xlabel = s; f.label_fs = fontsize


def ylabel(s, fontsize=12, **kw):
    f = _cur(); f.ylabel = s; f.label_fs = fontsize


def title(s, fontsize=12, **kw):
    f = _cur(); f.title = s; f.label_fs = fontsize


def close(*a, **kw):
    global _fig
    _fig = None


def _marker(d, x, y, m, col, r=6):
    if m == ".":
        d.ellipse([x - 3, y - 3, x + 3, y + 3], fill=col)
    elif m == "x":
        d.line([x - r, y - r, x + r, y + r], fill=col, width=2); d.line([x - r, y + r, x + r, y - r], fill=col, width=2)
    elif m == "+":
        d.line([x - r, y, x + r, y], fill=col, width=2); d.line([x, y - r, x, y + r], fill=col, width=2)
    elif m == ">":
        d.polygon([(x - r, y - r), (x + r, y), (x - r, y + r)], fill=col)


def savefig(fname, format="png", dpi=100, **kw):
    f = _cur()
    W, H = int(f.figsize[0] * dpi), int(f.figsize[1] * dpi)
    img = Image.new("RGB", (W, H), (255, 255, 255))
    d = ImageDraw.Draw(img)
    L, R, T, B = int(0.07 * W), int(0.98 * W), int(0.06 * H), int(0.90 * H)
    npts = max((len(s[0]) for s in f.ax.series), default=1)
    x0, x1 = -0.01 * max(1, npts - 1), (npts - 1) * 1.01 if npts > 1 else 1.0
    ylo = f.ylim[0] if f.ylim and f.ylim[0] is not None else 0.0
    yhi = f.ylim[1] if f.ylim and f.ylim[1] is not None else f.data_ymax() * 1.05
    if yhi <= ylo:
        yhi = ylo + 1.0
    px = lambda x: L + (x - x0) / (x1 - x0) * (R - L)
    py = lambda y: B - (y - ylo) / (yhi - ylo) * (B - T)
    d.rectangle([L, T, R, B], outline=(0, 0, 0), width=2)
    tf = _font(f.xticks[2] * 0.6)
    for loc, lab in zip(*f.xticks[:2]):
        x = px(loc)
        d.line([x, B, x, B + 10], fill=(0, 0, 0), width=2)
        d.text((x, B + 14), lab, fill=(0, 0, 0), font=tf, anchor="ma")
    yl = f.yticks[0] if f.yticks[0] is not None else list(f.auto_yticks())
    for v in yl:
        if ylo <= v <= yhi:
            y = py(v)
            d.line([L - 10, y, L, y], fill=(0, 0, 0), width=2)
            d.text((L - 14, y), f"{v:g}", fill=(0, 0, 0), font=tf, anchor="rm")
    for y, col, ls in f.ax.hlines:
        if ylo <= y <= yhi:
            yy = py(y)
            for xs in range(L, R, 24):
                d.line([xs, yy, min(xs + (12 if ls == "--" else 24), R), yy], fill=col, width=2)
    for xs, ys, col, m, _ in f.ax.series:
        pts = [(px(x), py(min(max(y, ylo), yhi))) for x, y in zip(xs, ys)]
        if len(pts) > 1:
            d.line(pts, fill=col, width=3)
        step = max(1, len(pts) // 400)
        for p in pts[::step]:
            _marker(d, p[0], p[1], m, col)
    for x, y, s, fs in f.ax.texts:
        d.text((px(x) + 6, py(min(max(y, ylo), yhi)) - 4), s, fill=(0, 0, 0), font=_font(fs * 0.7), anchor="ls")
    lf = _font(f.label_fs * 0.7)
    d.text(((L + R) / 2, H - 0.035 * H), f.xlabel, fill=(0, 0, 0), font=lf, anchor="mm")
    d.text(((L + R) / 2, 0.03 * H), f.title, fill=(0, 0, 0), font=lf, anchor="mm")
    if f.ylabel:
        tw = int(d.textlength(f.ylabel, font=lf)) + 8
        strip = Image.new("RGB", (tw, 40), (255, 255, 255))
        ImageDraw.Draw(strip).text((4, 4), f.ylabel, fill=(0, 0, 0), font=lf)
        strip = strip.rotate(90, expand=True)
        img.paste(strip, (8, (T + B) // 2 - strip.size[1] // 2))
    if f.legend:
        gf = _font(f.legend * 0.7)
        x, y = L + 20, T + 16
        for i, (_, _, col, m, lab) in enumerate(f.ax.series):
            cx = x + (i % 2) * 420
            cy = y + (i // 2) * 36
            d.line([cx, cy + 10, cx + 50, cy + 10], fill=col, width=4)
            _marker(d, cx + 25, cy + 10, m, col, r=8)
            d.text((cx + 60, cy + 10), lab or "", fill=(0, 0, 0), font=gf, anchor="lm")
    img.save(fname, format=format.upper())
