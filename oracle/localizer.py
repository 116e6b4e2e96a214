"""Restatement of Localizer.lua.  Test infrastructure only.

The reference builds its layer list by walking nngraph nodes (Localizer.lua:8-38).  nngraph is not
available, so the list is derived from the model description tables instead: the walk visits, in forward
order, every leaf module that has kW and kH -- i.e. the SpatialConvolution and SpatialMaxPooling modules
(model_utilities.lua:7-35); PReLU / SpatialDropout have no kW and are skipped (Localizer.lua:31).
"""
import math

from .rect import Rect


def trunk_layer_info(layers, upto_block):
    """Leaf modules of conv blocks 1..upto_block (model_utilities.lua:17-25)."""
    info = []
    for l in layers[:upto_block]:
        for _ in range(l["conv_steps"]):
            info.append(dict(kW=l["kW"], kH=l["kH"], dW=1, dH=1, padW=l["padW"], padH=l["padH"]))
        # nn.SpatialMaxPooling(2, 2, 2, 2):ceil() -- padW/padH default 0 (Localizer.lua:32)
        info.append(dict(kW=2, kH=2, dW=2, dH=2, padW=0, padH=0))
    return info


def head_layer_info(layers, anchor_net):
    """Layers seen from pnet.outnode.children[i], i<=#anchor_nets (model_utilities.lua:29-35,51-54)."""
    info = trunk_layer_info(layers, anchor_net["input"])
    k = anchor_net["kW"]
    info.append(dict(kW=k, kH=k, dW=1, dH=1, padW=0, padH=0))
    info.append(dict(kW=1, kH=1, dW=1, dH=1, padW=0, padH=0))
    return info


class Localizer:
    def __init__(self, layer_info):
        self.layers = list(layer_info)

    def inputToFeatureRect(self, rect, layer_index=None):  # Localizer.lua:41-67
        layer_index = layer_index or len(self.layers)
        rect = Rect(rect.minX, rect.minY, rect.maxX, rect.maxY)
        for l in self.layers[:layer_index]:
            if l["dW"] < l["kW"]:
                rect = rect.inflate(l["kW"] - l["dW"], l["kH"] - l["dH"])
            rect = rect.offset(l["padW"], l["padH"])
            rect.minX = rect.minX / l["dH"]  # sic: dH for X (Localizer.lua:52)
            rect.minY = rect.minY / l["dH"]
            # Lua's % on doubles is a - floor(a/b)*b, same sign convention as Python's float %
            if (rect.maxX - l["kW"]) % l["dW"] == 0:
                rect.maxX = max((rect.maxX - l["kW"]) / l["dW"] + 1, rect.minX + 1)
            else:
                rect.maxX = max(math.ceil((rect.maxX - l["kW"]) / l["dW"]) + 1, rect.minX + 1)
            if (rect.maxY - l["kH"]) % l["dH"] == 0:
                rect.maxY = max((rect.maxY - l["kH"]) / l["dW"] + 1, rect.minY + 1)  # sic: / dW (Localizer.lua:60)
            else:
                rect.maxY = max(math.ceil((rect.maxY - l["kH"]) / l["dH"]) + 1, rect.minY + 1)
        return rect.snapToInt()

    def featureToInputRect(self, minX, minY, maxX, maxY, layer_index=None):  # Localizer.lua:69-79
        layer_index = layer_index or len(self.layers)
        for l in reversed(self.layers[:layer_index]):
            minX = minX * l["dW"] - l["padW"]
            minY = minY * l["dH"] - l["padW"]  # sic: padW (Localizer.lua:74)
            maxX = maxX * l["dW"] - l["padH"] + l["kW"] - l["dW"]  # sic: padH (Localizer.lua:75)
            maxY = maxY * l["dH"] - l["padH"] + l["kH"] - l["dH"]
        return Rect(minX, minY, maxX, maxY)
