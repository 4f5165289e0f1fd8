"""Extracts the reference crate's PUBLIC surface as text (run in the build container, where /root/reference exists):
the items re-exported by src/lib.rs:45-49 with their pub fn signatures, enum variants and trait methods.
Writes tests/golden/rust_public_api.json; tests/test_rust_shim.py checks rust/src/lib.rs against it.

    python tests/golden/make_rust_api.py
"""
import json
import os
import re

REF = "/root/reference/src"


def _norm(sig):
    return re.sub(r"\s+", " ", sig).strip().rstrip("{").strip().replace("( ", "(").replace(" )", ")").replace(",)", ")")


def _blocks(src, header_re):
    """Bodies of every item whose header matches header_re (brace matched)."""
    out = []
    for m in re.finditer(header_re, src):
        i = src.index("{", m.end() - 1)
        depth, j = 0, i
        while True:
            if src[j] == "{":
                depth += 1
            elif src[j] == "}":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        out.append((m.group(0), src[i + 1:j]))
    return out


def _pub_fns(body):
    sigs = []
    for m in re.finditer(r"^\s*pub fn (?:[^;{\[]|\[[^\]]*\])*", body, re.M):
        sigs.append(_norm(m.group(0)))
    return sigs


def _trait_fns(body):
    return [_norm(m.group(0)) for m in re.finditer(r"^\s*fn (?:[^;{\[]|\[[^\]]*\])*", body, re.M)]


def _variants(body):
    """Variant names of an enum body (attributes, doc comments and payloads stripped)."""
    body = re.sub(r"//[^\n]*", "", body)
    body = re.sub(r"#\[[^\]]*\]", "", body)
    names, depth, tok = [], 0, ""
    for ch in body:
        if ch in "({":
            depth += 1
        elif ch in ")}":
            depth -= 1
        elif ch == "," and depth == 0:
            names.append(tok)
            tok = ""
            continue
        if depth == 0 and ch not in ")}":
            tok += ch
    names.append(tok)
    out = []
    for n in names:
        n = n.split("=")[0].strip()
        if n:
            out.append(re.match(r"\w+", n).group(0))
    return out


def extract():
    api = {}
    lib = open(os.path.join(REF, "lib.rs")).read()
    exports = []
    for m in re.finditer(r"^pub use (\w+)::\{?([^;}]*)\}?;", lib, re.M):
        if "benchmark" in lib[max(0, m.start() - 120):m.start()]:
            continue
        exports += [x.strip() for x in m.group(2).split(",") if x.strip()]
    api["exports"] = sorted(exports)

    enc = open(os.path.join(REF, "encoder.rs")).read()
    api["Encoder"] = _pub_fns(_blocks(enc, r"impl<W: JfifWrite> Encoder<W> \{")[0][1])
    api["Encoder<BufWriter<File>>"] = _pub_fns(_blocks(enc, r"impl Encoder<BufWriter<File>> \{")[0][1])
    api["SamplingFactor::fns"] = [s for s in _pub_fns(_blocks(enc, r"impl SamplingFactor \{")[0][1])]
    for name in ("JpegColorType", "ColorType", "SamplingFactor"):
        api[name] = _variants(_blocks(enc, r"pub enum %s \{" % name)[0][1])

    img = open(os.path.join(REF, "image_buffer.rs")).read()
    api["image_buffer::fns"] = [_norm(m.group(0)) for m in re.finditer(r"^pub fn [^{]*", img, re.M)]
    api["ImageBuffer"] = _trait_fns(_blocks(img, r"pub trait ImageBuffer \{")[0][1])

    wr = open(os.path.join(REF, "writer.rs")).read()
    api["JfifWrite"] = _trait_fns(_blocks(wr, r"pub trait JfifWrite \{")[0][1])
    api["PixelDensityUnit"] = _variants(_blocks(wr, r"pub enum PixelDensityUnit \{")[0][1])
    api["PixelDensity::fields"] = [_norm(m.group(0)) for m in re.finditer(r"^\s*pub \w+: [^\n]*?(?=,?\s*$)", _blocks(wr, r"pub struct PixelDensity \{")[0][1], re.M)]
    api["PixelDensity::fns"] = _pub_fns(_blocks(wr, r"impl PixelDensity \{")[0][1])

    q = open(os.path.join(REF, "quantization.rs")).read()
    api["QuantizationTableType"] = _variants(_blocks(q, r"pub enum QuantizationTableType \{")[0][1])

    err = open(os.path.join(REF, "error.rs")).read()
    api["EncodingError"] = _variants(_blocks(err, r"pub enum EncodingError \{")[0][1])
    api["EncodingError::impls"] = sorted(_norm(m.group(1)) for m in re.finditer(r"^impl (\S+(?: for)? ?\S*) for EncodingError", err, re.M))
    api["EncodingError::display"] = re.findall(r'"([^"]*\{\}[^"]*)"', err)
    return api


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rust_public_api.json")
    json.dump(extract(), open(out, "w"), indent=1, sort_keys=True)
    print("wrote", out)
