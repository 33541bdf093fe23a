"""Darknet .cfg emitters for the four model definitions the reference ships in config/
(yolov3, yolov3-tiny, yolov4, yolov4-tiny).  The emitted files contain only the keys the hot path
reads (yolo3/utils/parse_config.py:1-19 keeps every key as a string; yolo3/models/models.py:25-102
reads type, batch_normalize, filters, size, stride, activation, layers, groups, group_id, from, mask,
anchors, classes and the [net] width/height/channels).  Training-only keys (learning rate, jitter,
scale_x_y, iou_loss ...) are ignored by the reference and therefore not emitted.

`python -m yolo_deepsort_b200.cfgs <outdir>` writes config/*.cfg.
"""
import os
import sys

COCO_V3_ANCHORS = "10,13,  16,30,  33,23,  30,61,  62,45,  59,119,  116,90,  156,198,  373,326"
COCO_V4_ANCHORS = "12, 16, 19, 36, 40, 28, 36, 75, 76, 55, 72, 146, 142, 110, 192, 243, 459, 401"
TINY_ANCHORS = "10,14,  23,27,  37,58,  81,82,  135,169,  344,319"


class _Cfg:
    def __init__(self, width, height):
        self.out = [f"[net]\nwidth={width}\nheight={height}\nchannels=3\n"]

    def conv(self, filters, size, stride=1, act="leaky", bn=True):
        s = "[convolutional]\n"
        if bn:
            s += "batch_normalize=1\n"
        s += f"filters={filters}\nsize={size}\nstride={stride}\npad=1\nactivation={act}\n"
        self.out.append(s)

    def shortcut(self, frm=-3):
        self.out.append(f"[shortcut]\nfrom={frm}\nactivation=linear\n")

    def route(self, layers, groups=None, group_id=None):
        s = f"[route]\nlayers={layers}\n"
        if groups is not None:
            s += f"groups={groups}\ngroup_id={group_id}\n"
        self.out.append(s)

    def maxpool(self, size, stride):
        self.out.append(f"[maxpool]\nsize={size}\nstride={stride}\n")

    def upsample(self, stride=2):
        self.out.append(f"[upsample]\nstride={stride}\n")

    def yolo(self, mask, anchors, num, classes=80):
        self.out.append(f"[yolo]\nmask={mask}\nanchors={anchors}\nclasses={classes}\nnum={num}\n")

    def text(self):
        return "\n".join(self.out)


def yolov3(width=416, height=416):
    c = _Cfg(width, height)
    c.conv(32, 3)
    for filters, nres in ((64, 1), (128, 2), (256, 8), (512, 8), (1024, 4)):
        c.conv(filters, 3, 2)
        for _ in range(nres):
            c.conv(filters // 2, 1)
            c.conv(filters, 3)
            c.shortcut(-3)
    for i, (f, mask, lateral) in enumerate(((512, "6,7,8", None), (256, "3,4,5", 61), (128, "0,1,2", 36))):
        if lateral is not None:
            c.route("-4")
            c.conv(f, 1)
            c.upsample(2)
            c.route(f"-1, {lateral}")
        for _ in range(3):
            c.conv(f, 1)
            c.conv(f * 2, 3)
        c.conv(255, 1, act="linear", bn=False)
        c.yolo(mask, COCO_V3_ANCHORS, 9)
    return c.text()


def yolov3_tiny(width=416, height=416):
    c = _Cfg(width, height)
    for f in (16, 32, 64, 128, 256):
        c.conv(f, 3)
        c.maxpool(2, 2)
    c.conv(512, 3)
    c.maxpool(2, 1)
    c.conv(1024, 3)
    c.conv(256, 1)
    c.conv(512, 3)
    c.conv(255, 1, act="linear", bn=False)
    c.yolo("3,4,5", TINY_ANCHORS, 6)
    c.route("-4")
    c.conv(128, 1)
    c.upsample(2)
    c.route("-1, 8")
    c.conv(256, 3)
    c.conv(255, 1, act="linear", bn=False)
    c.yolo("1,2,3", TINY_ANCHORS, 6)
    return c.text()


def yolov4(width=608, height=608):
    c = _Cfg(width, height)
    m = "mish"
    c.conv(32, 3, act=m)
    # CSP stages: (downsample filters, split filters, residual inner filters, n blocks, merge filters)
    for down, split, inner, n, merge in ((64, 64, 32, 1, 64), (128, 64, 64, 2, 128), (256, 128, 128, 8, 256),
                                         (512, 256, 256, 8, 512), (1024, 512, 512, 4, 1024)):
        c.conv(down, 3, 2, act=m)
        c.conv(split, 1, act=m)
        c.route("-2")
        c.conv(split, 1, act=m)
        for _ in range(n):
            c.conv(inner, 1, act=m)
            c.conv(split, 3, act=m)
            c.shortcut(-3)
        c.conv(split, 1, act=m)
        c.route(f"-1,-{3 * n + 4}")
        c.conv(merge, 1, act=m)
    # SPP neck
    c.conv(512, 1); c.conv(1024, 3); c.conv(512, 1)
    c.maxpool(5, 1); c.route("-2"); c.maxpool(9, 1); c.route("-4"); c.maxpool(13, 1)
    c.route("-1,-3,-5,-6")
    c.conv(512, 1); c.conv(1024, 3); c.conv(512, 1)
    # top-down
    for f, lateral in ((256, 85), (128, 54)):
        c.conv(f, 1)
        c.upsample(2)
        c.route(str(lateral))
        c.conv(f, 1)
        c.route("-1, -3")
        for k in range(5):
            c.conv(f if k % 2 == 0 else f * 2, 1 if k % 2 == 0 else 3)
    # heads, bottom-up
    c.conv(256, 3)
    c.conv(255, 1, act="linear", bn=False)
    c.yolo("0,1,2", COCO_V4_ANCHORS, 9)
    for f, back, mask in ((256, -16, "3,4,5"), (512, -37, "6,7,8")):
        c.route("-4")
        c.conv(f, 3, 2)
        c.route(f"-1, {back}")
        for k in range(5):
            c.conv(f if k % 2 == 0 else f * 2, 1 if k % 2 == 0 else 3)
        c.conv(f * 2, 3)
        c.conv(255, 1, act="linear", bn=False)
        c.yolo(mask, COCO_V4_ANCHORS, 9)
    return c.text()


def yolov4_tiny(width=416, height=416):
    c = _Cfg(width, height)
    c.conv(32, 3, 2)
    c.conv(64, 3, 2)
    for f in (64, 128, 256):
        c.conv(f, 3)
        c.route("-1", groups=2, group_id=1)
        c.conv(f // 2, 3)
        c.conv(f // 2, 3)
        c.route("-1,-2")
        c.conv(f, 1)
        c.route("-6,-1")
        c.maxpool(2, 2)
    c.conv(512, 3)
    c.conv(256, 1)
    c.conv(512, 3)
    c.conv(255, 1, act="linear", bn=False)
    c.yolo("3,4,5", TINY_ANCHORS, 6)
    c.route("-4")
    c.conv(128, 1)
    c.upsample(2)
    c.route("-1, 23")
    c.conv(256, 3)
    c.conv(255, 1, act="linear", bn=False)
    c.yolo("1,2,3", TINY_ANCHORS, 6)
    return c.text()


ALL = {"yolov3": yolov3, "yolov3-tiny": yolov3_tiny, "yolov4": yolov4, "yolov4-tiny": yolov4_tiny}


def write_all(outdir):
    os.makedirs(outdir, exist_ok=True)
    for name, fn in ALL.items():
        with open(os.path.join(outdir, name + ".cfg"), "w") as f:
            f.write(fn())


if __name__ == "__main__":
    write_all(sys.argv[1] if len(sys.argv) > 1 else "config")
