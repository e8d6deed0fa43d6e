"""Stand-in for the `xmltodict` package (absent from this image; no network).

TEST INFRASTRUCTURE ONLY -- used solely by tests/golden/make_golden.py to import the
Python reference from /root/reference inside this container.  It follows xmltodict's
documented mapping for the subset the reference's XML files use:
element with children -> dict, leaf -> stripped text (None when empty), repeated sibling
tags -> list.  All leaves stay strings; the reference casts them itself.
"""
import xml.etree.ElementTree as ET


def _conv(el):
    kids = list(el)
    if not kids:
        t = el.text.strip() if el.text and el.text.strip() else None
        return t
    out = {}
    for k in kids:
        v = _conv(k)
        if k.tag in out:
            if not isinstance(out[k.tag], list):
                out[k.tag] = [out[k.tag]]
            out[k.tag].append(v)
        else:
            out[k.tag] = v
    return out


def parse(xml_text, *a, **kw):
    root = ET.fromstring(xml_text)
    return {root.tag: _conv(root)}
