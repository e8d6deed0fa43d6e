"""XML -> nested dict of strings, the mapping the reference gets from xmltodict.parse
(BaseClass/CalMod.py:117-124): element with children -> dict, leaf -> stripped text (None if empty),
repeated sibling tags -> list.  Consumers cast and default exactly like the reference does
(None2Value, CalMod.py:56-61)."""
import xml.etree.ElementTree as ET


def _conv(el):
    kids = list(el)
    if not kids:
        return el.text.strip() if el.text and el.text.strip() else None
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


def XML2Dict(file_path):
    with open(file_path, "r") as f:
        root = ET.fromstring(f.read())
    return {root.tag: _conv(root)}


def None2Value(value1, value2=None):
    return value2 if value1 is None else value1


def epsilon_annealing(epoch, min_eps, max_eps_episode):
    """simulator.epsilon_annealing (simulator.py:141-145)."""
    slope = (min_eps - 1.0) / (max_eps_episode + 0.1)
    return max(slope * epoch + 1.0, min_eps)
