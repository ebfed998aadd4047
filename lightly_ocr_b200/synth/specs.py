"""Layer tables of the reference graphs (state-dict keys and shapes).

CRAFT : reference ocr/model.py:9-37 (VGG_UNet), ocr/modules/vgg_bn.py:23-55 (UpConv, vgg16_bn slices).
CRNN  : reference ocr/model.py:64-101 (CRNNet), ocr/modules/TPS_STN.py:32-68, ocr/modules/resnet50v1.py:51-99,
        ocr/modules/biLSTM.py:9-20, ocr/modules/attention.py:8-15,64-72.
Each conv entry: (state-dict prefix, Cin, Cout, kernel, bn prefix or None, has_bias).
"""

# torchvision vgg16_bn feature indices as re-registered by the reference's slices (vgg_bn.py:44-51)
CRAFT_CONVS = [
    ("basenet.slice1.0", 3, 64, 3, "basenet.slice1.1", True),
    ("basenet.slice1.3", 64, 64, 3, "basenet.slice1.4", True),
    ("basenet.slice1.7", 64, 128, 3, "basenet.slice1.8", True),
    ("basenet.slice1.10", 128, 128, 3, "basenet.slice1.11", True),
    ("basenet.slice2.14", 128, 256, 3, "basenet.slice2.15", True),
    ("basenet.slice2.17", 256, 256, 3, "basenet.slice2.18", True),
    ("basenet.slice3.20", 256, 256, 3, "basenet.slice3.21", True),
    ("basenet.slice3.24", 256, 512, 3, "basenet.slice3.25", True),
    ("basenet.slice3.27", 512, 512, 3, "basenet.slice3.28", True),
    ("basenet.slice4.30", 512, 512, 3, "basenet.slice4.31", True),
    ("basenet.slice4.34", 512, 512, 3, "basenet.slice4.35", True),
    ("basenet.slice4.37", 512, 512, 3, "basenet.slice4.38", True),
    ("basenet.slice5.1", 512, 1024, 3, None, True),     # dilation 6, padding 6 (vgg_bn.py:54)
    ("basenet.slice5.2", 1024, 1024, 1, None, True),
    ("upconv1.conv.0", 1536, 512, 1, "upconv1.conv.1", True),
    ("upconv1.conv.3", 512, 256, 3, "upconv1.conv.4", True),
    ("upconv2.conv.0", 768, 256, 1, "upconv2.conv.1", True),
    ("upconv2.conv.3", 256, 128, 3, "upconv2.conv.4", True),
    ("upconv3.conv.0", 384, 128, 1, "upconv3.conv.1", True),
    ("upconv3.conv.3", 128, 64, 3, "upconv3.conv.4", True),
    ("upconv4.conv.0", 192, 64, 1, "upconv4.conv.1", True),
    ("upconv4.conv.3", 64, 32, 3, "upconv4.conv.4", True),
    ("conv_cls.0", 32, 32, 3, None, True),
    ("conv_cls.2", 32, 32, 3, None, True),
    ("conv_cls.4", 32, 16, 3, None, True),
    ("conv_cls.6", 16, 16, 1, None, True),
    ("conv_cls.8", 16, 2, 1, None, True),
]

LOC = "Transformation.LocalizationNetwork."
FE = "FeatureExtraction.ConvNet."

# (prefix, Cin, Cout, (kh, kw), bn prefix)  - all CRNN convs are bias-free and followed by BN
CRNN_LOC_CONVS = [
    (LOC + "conv.0", 1, 64, (3, 3), LOC + "conv.1"),
    (LOC + "conv.4", 64, 128, (3, 3), LOC + "conv.5"),
    (LOC + "conv.8", 128, 256, (3, 3), LOC + "conv.9"),
    (LOC + "conv.12", 256, 512, (3, 3), LOC + "conv.13"),
]


def _blocks(layer, n, cin, planes):
    out = []
    for i in range(n):
        p = "%slayer%d.%d." % (FE, layer, i)
        c_in = cin if i == 0 else planes
        out.append((p + "conv1", c_in, planes, (3, 3), p + "bn1"))
        out.append((p + "conv2", planes, planes, (3, 3), p + "bn2"))
        if i == 0 and c_in != planes:
            out.append((p + "downsample.0", c_in, planes, (1, 1), p + "downsample.1"))
    return out


# resnet50v1.py:55-82 with BasicBlock counts [1, 2, 5, 3] (resnet50v1.py:10)
CRNN_FE_CONVS = (
    [(FE + "conv0_1", 1, 32, (3, 3), FE + "bn0_1"), (FE + "conv0_2", 32, 64, (3, 3), FE + "bn0_2")]
    + _blocks(1, 1, 64, 128) + [(FE + "conv1", 128, 128, (3, 3), FE + "bn1")]
    + _blocks(2, 2, 128, 256) + [(FE + "conv2", 256, 256, (3, 3), FE + "bn2")]
    + _blocks(3, 5, 256, 512) + [(FE + "conv3", 512, 512, (3, 3), FE + "bn3")]
    + _blocks(4, 3, 512, 512)
    + [(FE + "conv4_1", 512, 512, (2, 2), FE + "bn4_1"), (FE + "conv4_2", 512, 512, (2, 2), FE + "bn4_2")]
)
RESNET_BLOCKS = {1: 1, 2: 2, 3: 5, 4: 3}

NUM_FIDUCIAL = 20
IMG_H, IMG_W = 32, 100
HIDDEN = 256
SEQ_T = 26
ALPHABET = "0123456789abcdefghijklmnopqrstuvwxyz"
