"""Host-side mirrors of the reference drivers: ``ImageDetector`` (yolo3/detect/img_detect.py:37-153, single-window
path) and ``VideoDetector`` (yolo3/detect/video_detect.py:39-208).  Host plumbing only (capture, frame skipping,
drawing, FPS read-out); detection, ReID and tracking run in libydst.
"""
import logging
import time
from functools import reduce

import cv2
import numpy as np
import torch

from .darknet import p1p2Toxywh, resize_boxes, soft_non_max_suppression
from .deepsort import DeepSort
from .pipeline import FramePipeline
from .reid import Extractor


def load_classes(path):
    with open(path, "r", encoding="utf-8") as fp:
        return fp.read().split("\n")[:-1]


class ImageDetector:
    def __init__(self, model, class_path, thickness=2, thres=0.5, nms_thres=0.4, win_size=None, overlap=0.15, half=False):
        self.model = model
        self.model.eval()
        self.device = next(self.model.parameters()).device
        if half:
            self.model.half()
        self.classes = load_classes(class_path) if class_path else []
        self.num_classes = len(self.classes)
        self.thickness, self.thres, self.nms_thres, self.half = thickness, thres, nms_thres, half
        if win_size is not None:
            raise NotImplementedError("sliding-window mode (win_size) is not part of the accelerated path yet")
        self.win_size, self.overlap = win_size, overlap

    def detect(self, img):
        """img: (h,w,3) uint8 RGB.  Returns (n,6) float32 tensor [x1,y1,x2,y2,conf,cls] in image pixels, or None."""
        h, w, _ = img.shape
        H, W = self.model.img_size
        image = img if (h, w) == (H, W) else cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR)
        frame = torch.from_numpy(np.ascontiguousarray(image)).to(self.device)
        t0 = time.time()
        pred = self.model.forward_frame(frame)
        dets = soft_non_max_suppression(pred, self.thres, self.nms_thres)[0]
        if dets is not None:
            dets = resize_boxes(dets, self.model.img_size, (h, w))
        logging.info("\t Inference time: %.6f s" % (time.time() - t0))
        return dets


def _color(i):
    rng = np.random.default_rng(int(i) * 7919 + 13)
    return tuple(int(c) for c in rng.integers(64, 256, 3))


class VideoDetector:
    def __init__(self, model, class_path, thickness=2, font_path=None, font_size=10, thres=0.7, nms_thres=0.4, skip_frames=-1,
                 fourcc=cv2.VideoWriter_fourcc('m', 'p', '4', 'v'), class_mask=None, win_size=None, overlap=0.15, tracker=None,
                 action_id=None, half=False):
        self.thickness, self.skip_frames, self.class_mask, self.fourcc = thickness, skip_frames, class_mask, fourcc
        self.image_detector = ImageDetector(model, class_path, thickness=thickness, thres=thres, nms_thres=nms_thres,
                                            win_size=win_size, overlap=overlap, half=half)
        self.classes = self.image_detector.classes
        self.tracker, self.action_id = tracker, action_id
        self._pipeline = None
        if isinstance(tracker, DeepSort) and isinstance(tracker.extractor, Extractor):
            self._pipeline = FramePipeline(model, tracker, thres, nms_thres, class_mask)

    def _track(self, frame):
        """One detection step; returns `hold_detections` exactly as the reference loop would set it
        (yolo3/detect/video_detect.py:134-157)."""
        if self._pipeline is not None:
            self._pipeline.submit(frame, want_dets=False)        # any frame size: the resize runs on the device
            tracks, dets = self._pipeline.collect(want_dets=False)
            return tracks
        detections = self.image_detector.detect(frame)
        if detections is not None and self.tracker is not None:
            boxs = p1p2Toxywh(detections[:, :4])
            class_ids, confidences = detections[:, -1], detections[:, 4]
            if self.class_mask is not None:
                mask = reduce(lambda a, b: a | b, [class_ids == mid for mid in self.class_mask])
                boxs, confidences, class_ids = boxs[mask], confidences[mask], class_ids[mask]
            detections = self.tracker.update(boxs.float(), confidences, frame, class_ids)
        return detections

    def _draw(self, frame, hold):
        img = frame.copy()
        if hold is None or len(hold) == 0:
            return img
        tracked = self.tracker is not None
        for row in hold:
            x1, y1, x2, y2 = (int(v) for v in row[:4])
            if tracked:
                tid, cid = int(row[4]), int(row[5])
                label = f"{self.classes[cid] if 0 <= cid < len(self.classes) else cid} #{tid}"
                col = _color(tid)
            else:
                cid = int(row[5])
                label = f"{self.classes[cid] if 0 <= cid < len(self.classes) else cid} {float(row[4]):.2f}"
                col = _color(cid)
            cv2.rectangle(img, (x1, y1), (x2, y2), col, self.thickness)
            cv2.putText(img, label, (x1, max(y1 - 3, 10)), cv2.FONT_HERSHEY_SIMPLEX, 0.5, col, 1)
        return img

    def detect(self, video_path, output_path=None, skip_secs=0, real_show=False, show_fps=True):
        logging.info("Detect video: " + str(video_path))
        vid = cv2.VideoCapture(video_path)
        if not vid.isOpened():
            raise IOError("Couldn't open webcam or video")
        video_fps = int(vid.get(cv2.CAP_PROP_FPS))
        video_size = (int(vid.get(cv2.CAP_PROP_FRAME_WIDTH)), int(vid.get(cv2.CAP_PROP_FRAME_HEIGHT)))
        total_frames = int(vid.get(cv2.CAP_PROP_FRAME_COUNT))
        if skip_secs > total_frames:
            print("Can't skip over total video!")
        else:
            vid.set(cv2.CAP_PROP_POS_FRAMES, int(skip_secs) * video_fps)
        out = cv2.VideoWriter(output_path, self.fourcc, video_fps, video_size) if output_path is not None else None
        if real_show:
            cv2.namedWindow("result", cv2.WINDOW_NORMAL)
            cv2.resizeWindow("result", 960, 540)
        accum_time, curr_fps, fps, prev_time = 0, 0, "FPS: ??", time.time()
        hold_detections, actions, frames = None, [], 0
        H, W = self.image_detector.model.img_size

        def read_rgb():
            ok, bgr = vid.read()
            return cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB) if ok and bgr is not None else None

        # One frame of look-ahead (the reference's reader thread decodes ahead as well, video_detect.py:86,112): when every
        # frame is a detection frame and the fused pipeline applies, the detector of frame t+1 is submitted before the
        # ReID + association of frame t is collected, so the two halves overlap on the GPU.  Results are unchanged.
        lookahead = self._pipeline is not None and self.skip_frames in (-1, 1) and self.action_id is None
        nxt = read_rgb()
        submitted = False
        try:
            while nxt is not None:
                frame, nxt = nxt, read_rgb()
                if lookahead:
                    if not submitted:
                        self._pipeline.submit(frame, want_dets=False)
                    submitted = nxt is not None
                    if submitted:
                        self._pipeline.submit(nxt, want_dets=False)
                    detections, _ = self._pipeline.collect(want_dets=False)
                    actions = []
                    hold_detections = detections
                    frames = 0
                elif frames % self.skip_frames == 0:
                    detections = self._track(frame)
                    if detections is not None and self.tracker is not None and self.action_id is not None:
                        actions = self.action_id.update(detections)
                    else:
                        actions = []
                    hold_detections = detections
                    frames = 0
                else:
                    actions = []
                hold = hold_detections.cpu().numpy() if isinstance(hold_detections, torch.Tensor) else hold_detections
                result = cv2.cvtColor(self._draw(frame, hold), cv2.COLOR_RGB2BGR)
                frames += 1
                curr_time = time.time()
                accum_time += curr_time - prev_time
                prev_time = curr_time
                curr_fps += 1
                if accum_time > 1:
                    accum_time -= 1
                    fps = "FPS: " + str(curr_fps)
                    curr_fps = 0
                    print(fps)
                if show_fps:
                    cv2.putText(result, text=fps, org=(3, 15), fontFace=cv2.FONT_HERSHEY_SIMPLEX, fontScale=0.6, color=(255, 0, 0),
                                thickness=self.thickness)
                if real_show:
                    cv2.imshow("result", result)
                if out is not None:
                    out.write(result)
                yield result, hold_detections, actions
                if real_show and cv2.waitKey(1) & 0xFF == ord('q'):
                    break
        finally:
            if self._pipeline is not None:
                self._pipeline.drain()
            vid.release()
            if out is not None:
                out.release()
            if real_show:
                cv2.destroyAllWindows()
