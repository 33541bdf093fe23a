"""Host-side mirrors of the reference drivers: ``ImageDetector`` (yolo3/detect/img_detect.py:37-153, single-window
and sliding-window paths) and ``VideoDetector`` (yolo3/detect/video_detect.py:39-208).  Host plumbing only (capture, frame skipping,
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
from .label_draw import LabelDrawer
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
        self.win_size, self.overlap = win_size, overlap

    def detect(self, img):
        """img: (h,w,3) uint8 RGB.  Returns (n,6) float32 tensor [x1,y1,x2,y2,conf,cls] in image pixels, or None."""
        h, w, _ = img.shape
        H, W = self.model.img_size
        if self.win_size is not None:
            win_width, win_height = self.win_size
        if self.win_size is None or w < win_width and h < win_height:        # (the reference's condition, img_detect.py:68)
            image = img if (h, w) == (H, W) else cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR)
            frame = torch.from_numpy(np.ascontiguousarray(image)).to(self.device)
            t0 = time.time()
            pred = self.model.forward_frame(frame)
            dets = soft_non_max_suppression(pred, self.thres, self.nms_thres)[0]
            if dets is not None:
                dets = resize_boxes(dets, self.model.img_size, (h, w))
        else:
            t0 = time.time()
            dets = self._detect_windows(img, win_width, win_height)
        logging.info("\t Inference time: %.6f s" % (time.time() - t0))
        return dets

    def _detect_windows(self, img, win_width, win_height):
        """Sliding-window mode (yolo3/detect/img_detect.py:97-151): windows of win_size plus `overlap` of it to the right and
        below (x outer, y inner) are cut and cv2-exactly resized ON THE DEVICE, go through the detector as one batch (one
        forward with M = tiles x grid cells: the natural batch > 1 use of the tcgen05 convolutions), their boxes are mapped back
        to image pixels and the union is reduced by soft_non_max_suppression(merge=True, is_p1p2=True)."""
        from ._lib import check, lib, ptr, stream_ptr
        import ctypes
        h, w, _ = img.shape
        H, W = self.model.img_size
        ov_x, ov_y = int(win_width * self.overlap), int(win_height * self.overlap)
        rois = [(x, y, min(w, x + win_width + ov_x) - x, min(h, y + win_height + ov_y) - y)
                for x in range(0, w, win_width) for y in range(0, h, win_height)]
        T = len(rois)
        with torch.cuda.device(self.device):
            frame = torch.from_numpy(np.ascontiguousarray(img)).to(self.device)
            tiles = torch.empty((T, H, W, 3), dtype=torch.uint8, device=self.device)
            for t, (x, y, rw, rh) in enumerate(rois):
                check(lib().ydst_resize_u8_roi(ptr(frame), h, w, x, y, rw, rh, ptr(tiles[t]), H, W, 0, stream_ptr()))
            hdl = self.model.handle(T)
            rows, fields = self.model._shape(hdl)
            pred = torch.empty((T, rows, fields), dtype=torch.float32, device=self.device)
            check(lib().ydst_detector_forward_u8(hdl, ptr(tiles), ptr(pred), stream_ptr()))
            ratios = np.asarray([[rw / W, rh / H] for (_, _, rw, rh) in rois], np.float32)     # resize_boxes' python-float ratios
            offsets = np.asarray([[x, y] for (x, y, _, _) in rois], np.float32)
            check(lib().ydst_window_boxes(ptr(pred), T, rows, fields, ratios.ctypes.data, offsets.ctypes.data, stream_ptr()))
            return soft_non_max_suppression(pred.view(1, T * rows, fields), self.thres, self.nms_thres, merge=True, is_p1p2=True)[0]


class FrameReader:
    """The reference's reader thread (imutils.video.FileVideoStream with transform=BGR->RGB, yolo3/detect/video_detect.py:33-36,86,112):
    decodes and colour-converts up to 128 frames ahead of the loop on its own thread, so the decode overlaps the GPU work."""

    def __init__(self, vid, queue_size=128):
        import queue
        import threading
        self.vid, self.q, self.stopped = vid, queue.Queue(maxsize=queue_size), False
        self.thread = threading.Thread(target=self._run, daemon=True)

    def start(self):
        self.thread.start()
        return self

    def _run(self):
        while not self.stopped:
            ok, bgr = self.vid.read()
            if not ok or bgr is None:
                break
            self.q.put(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
        self.q.put(None)

    def read(self):
        return self.q.get()

    def stop(self):
        self.stopped = True
        try:
            while True:
                self.q.get_nowait()
        except Exception:
            pass
        self.thread.join(timeout=2)


class VideoDetector:
    def __init__(self, model, class_path, thickness=2, font_path=None, font_size=10, thres=0.7, nms_thres=0.4, skip_frames=-1,
                 fourcc=cv2.VideoWriter_fourcc('m', 'p', '4', 'v'), class_mask=None, win_size=None, overlap=0.15, tracker=None,
                 action_id=None, half=False):
        self.thickness, self.skip_frames, self.class_mask, self.fourcc = thickness, skip_frames, class_mask, fourcc
        self.image_detector = ImageDetector(model, class_path, thickness=thickness, thres=thres, nms_thres=nms_thres,
                                            win_size=win_size, overlap=overlap, half=half)
        self.classes = self.image_detector.classes
        self.label_drawer = LabelDrawer(self.classes, font_path, font_size, thickness, img_size=model.img_size)
        self.tracker, self.action_id = tracker, action_id
        self._pipeline = None
        # (the sliding-window mode goes through ImageDetector.detect + tracker.update, like the reference loop)
        if isinstance(tracker, DeepSort) and isinstance(tracker.extractor, Extractor) and win_size is None:
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
        """The overlay of the reference loop (yolo3/detect/video_detect.py:159-169): drawn INTO the frame, as there."""
        if hold is None:
            return frame
        if self.tracker is not None:
            image, _, _ = self.label_drawer.draw_labels_by_trackers(frame, hold, only_rect=False)
        else:
            image, _, _ = self.label_drawer.draw_labels(frame, hold, only_rect=False)
        return image

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

        reader = FrameReader(vid).start()
        read_rgb = reader.read

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
                hold = hold_detections.cpu().numpy() if isinstance(hold_detections, torch.Tensor) and self.tracker is not None \
                    else hold_detections
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
            reader.stop()
            vid.release()
            if out is not None:
                out.release()
            if real_show:
                cv2.destroyAllWindows()
