"""Overlay: mirror of ``LabelDrawer`` (yolo3/utils/label_draw.py:112-191) and its helpers ``draw_rects`` (:17-27),
``draw_rects_and_labels`` (:30-62) and ``draw_single_img`` (:65-100).  Host code, like the reference's: the very same cv2
calls with the same arguments, so the pixels are identical (tests/golden/overlay.npz holds the reference's output).  It runs
next to the hot path -- VideoDetector draws while the GPU works on the following frames."""
import logging

import cv2
import numpy as np


def draw_rects(img, dets, colors, thickness):
    for det in dets:
        x1, y1, x2, y2 = det[:4]
        cls = int(det[-1])
        cv2.rectangle(img, (int(x1), int(y1)), (int(x2), int(y2)), colors[cls], thickness)
    return img


def draw_rects_and_labels(img, dets, colors, labels, thickness, font_size, font=None):
    for i, det in enumerate(dets):
        x1, y1, x2, y2 = det[:4]
        cls = int(det[-1])
        c1, c2 = (int(x1), int(y1)), (int(x2), int(y2))
        cv2.rectangle(img, c1, c2, colors[cls], thickness)
        if font is not None:
            (font_w, font_h), _ = font.getTextSize(labels[i], font_size, -1)
            cv2.rectangle(img, (c1[0], max(0, c1[1] - 3 - font_size)), (c1[0] + font_w, max(c1[1], 3 + font_size)), colors[cls], -1)
            font.putText(img=img, text=labels[i], org=(c1[0], max(c1[1] - 3, font_size)), fontHeight=font_size, color=(0, 0, 0),
                         thickness=-1, line_type=cv2.LINE_4, bottomLeftOrigin=True)
        else:
            (font_w, font_h), _ = cv2.getTextSize(labels[i], cv2.FONT_HERSHEY_COMPLEX_SMALL, font_size, 1)
            cv2.rectangle(img, (c1[0], max(0, int(c1[1] - 3 - 18 * font_size))), (c1[0] + font_w, max(c1[1], int(3 + 18 * font_size))),
                          colors[cls], -1)
            cv2.putText(img, labels[i], (c1[0], max(c1[1] - 3, font_h)), cv2.FONT_HERSHEY_COMPLEX_SMALL, font_size, (0, 0, 0), 1)
    return img


def draw_single_img(img, detections, img_size, classes, colors, thickness, font, statistic=False, scaled=False, only_rect=False,
                    font_size=18):
    if detections is None:
        logging.debug("Nothing Detected.")
        return img, None, None
    detections = detections.cpu().float().numpy() if hasattr(detections, "cpu") else np.asarray(detections, np.float32)
    if only_rect:
        draw_rects(img, detections, colors, thickness)
    else:
        labels = []
        for d in detections:
            conf = d[-3] * d[-2] if len(d) == 7 else d[-2]
            labels.append(classes[int(d[-1])] + ' (' + str(round(conf * 100, 2)) + '%)')
        draw_rects_and_labels(img, detections, colors, labels, thickness, font_size, font)
    return img, None, None


class LabelDrawer:
    def __init__(self, classes, font_path, font_size, thickness, img_size, statistic=False, id2label=None):
        self.thickness, self.statistic, self.classes, self.img_size = thickness, statistic, classes, img_size
        self.font_size, self.id2label, self.font_path = font_size, id2label, font_path
        if font_path is not None:
            self.font = cv2.freetype.createFreeType2()
            self.font.loadFontData(fontFileName=font_path, id=0)
        else:
            self.font = None
        # one colour per class from numpy's legacy generator seeded with 1 (label_draw.py:140-145); a private RandomState yields
        # the same stream without disturbing the global one
        colors = (np.random.RandomState(1).rand(min(999, len(classes)), 3) * 255).astype(int)
        self.colors = [(int(c[0]), int(c[1]), int(c[2])) for c in colors]

    def clone(self):
        return LabelDrawer(self.classes, self.font_path, self.font_size, self.thickness, self.img_size, self.statistic, None)

    def draw_labels(self, img, detections, only_rect, scaled=True):
        return draw_single_img(img, detections, self.img_size, self.classes, self.colors, self.thickness, self.font,
                               statistic=self.statistic, scaled=scaled, only_rect=only_rect,
                               font_size=img.shape[0] / 1000. if self.font is None else self.font_size)

    def draw_labels_by_trackers(self, img, detections, only_rect):
        if only_rect:
            draw_rects(img, detections, self.colors, self.thickness)
        else:
            labels = []
            for d in detections:
                key = str(int(d[4]))
                labels.append(key + ":" + (self.id2label[key] if self.id2label is not None and key in self.id2label
                                           else self.classes[int(d[-1])]))
            draw_rects_and_labels(img, detections, self.colors, labels, self.thickness,
                                  img.shape[0] / 1000. if self.font is None else self.font_size, self.font)
        return img, None, None
