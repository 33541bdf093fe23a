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
    """Mirror of yolo3/detect/video_detect.py:39-208.  Same constructor keywords, same generator contract
    (`detect()` yields (bgr image with the overlay, held rows, actions) per frame, in order).  Two keywords of its own:
      micro_batch   consecutive detection frames per Darknet / ReID forward when reading a FILE (default 8; results per frame are
                    those of batch 1, tests/test_gpu_pipeline.py).  The reference's reader thread already decodes up to 128 frames
                    ahead of the loop (video_detect.py:86), so the look-ahead stays inside its design; a live source (camera index)
                    always runs batch 1 with one frame of look-ahead, where latency matters.
      draw_workers  overlay + colour conversion of up to this many frames run on worker threads (cv2 releases the GIL) while the
                    main thread collects the next results; frames are still yielded in order."""

    def __init__(self, model, class_path, thickness=2, font_path=None, font_size=10, thres=0.7, nms_thres=0.4, skip_frames=-1,
                 fourcc=cv2.VideoWriter_fourcc('m', 'p', '4', 'v'), class_mask=None, win_size=None, overlap=0.15, tracker=None,
                 action_id=None, half=False, micro_batch=8, draw_workers=4):
        self.thickness, self.skip_frames, self.class_mask, self.fourcc = thickness, skip_frames, class_mask, fourcc
        self.image_detector = ImageDetector(model, class_path, thickness=thickness, thres=thres, nms_thres=nms_thres,
                                            win_size=win_size, overlap=overlap, half=half)
        self.classes = self.image_detector.classes
        self.label_drawer = LabelDrawer(self.classes, font_path, font_size, thickness, img_size=model.img_size)
        self.tracker, self.action_id = tracker, action_id
        self.micro_batch, self.draw_workers = max(1, int(micro_batch)), max(1, int(draw_workers))
        self._model, self._thres, self._nms_thres = model, thres, nms_thres
        self._pipelines = {}
        # (the sliding-window mode goes through ImageDetector.detect + tracker.update, like the reference loop)
        self._fused = isinstance(tracker, DeepSort) and isinstance(tracker.extractor, Extractor) and win_size is None
        self._pipeline = self._pipeline_for(1) if self._fused else None

    def _pipeline_for(self, micro_batch):
        if micro_batch not in self._pipelines:
            self._pipelines[micro_batch] = FramePipeline(self._model, self.tracker, self._thres, self._nms_thres, self.class_mask,
                                                         micro_batch=micro_batch)
        return self._pipelines[micro_batch]

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
        """The overlay of the reference loop (yolo3/detect/video_detect.py:159-172): drawn INTO the frame, as there, then RGB -> BGR."""
        if hold is not None:
            hold = hold.cpu().numpy() if isinstance(hold, torch.Tensor) and self.tracker is not None else hold
            if self.tracker is not None:
                frame, _, _ = self.label_drawer.draw_labels_by_trackers(frame, hold, only_rect=False)
            else:
                frame, _, _ = self.label_drawer.draw_labels(frame, hold, only_rect=False)
        return cv2.cvtColor(frame, cv2.COLOR_RGB2BGR)

    def detect(self, video_path, output_path=None, skip_secs=0, real_show=False, show_fps=True):
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor
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
        state = {"accum": 0.0, "curr_fps": 0, "fps": "FPS: ??", "prev": time.time()}
        reader = FrameReader(vid).start()
        live = not isinstance(video_path, str)
        B = 1 if live else self.micro_batch
        pipe = self._pipeline_for(B) if self._fused else None
        # frames of look-ahead: the pipeline's three slots for a file, ONE frame for a live source
        max_ahead = 2 if live else 10 ** 9
        pool = ThreadPoolExecutor(max_workers=self.draw_workers)
        skip = self.skip_frames

        def finish(result):
            """Main-thread tail of one frame, in order: FPS read-out, show / write (yolo3/detect/video_detect.py:174-198)."""
            now = time.time()
            state["accum"] += now - state["prev"]
            state["prev"] = now
            state["curr_fps"] += 1
            if state["accum"] > 1:
                state["accum"] -= 1
                state["fps"] = "FPS: " + str(state["curr_fps"])
                state["curr_fps"] = 0
                print(state["fps"])
            if show_fps:
                cv2.putText(result, text=state["fps"], org=(3, 15), fontFace=cv2.FONT_HERSHEY_SIMPLEX, fontScale=0.6, color=(255, 0, 0),
                            thickness=self.thickness)
            if real_show:
                cv2.imshow("result", result)
            if out is not None:
                out.write(result)
            return result

        waiting, drawing = deque(), deque()                  # frames read but not collected | frames being drawn
        hold_detections, frames, eof = None, 0, False
        try:
            while True:
                # ---- read ahead: every frame the reference loop would run the detector on goes into the pipeline ----
                while not eof and len(waiting) < max_ahead and (pipe is None or pipe.can_submit() or not waiting):
                    frame = reader.read()
                    if frame is None:
                        eof = True
                        break
                    is_det = frames % skip == 0                # (python: x % -1 == 0 for every x, i.e. skip_frames=-1 detects on every frame)
                    if is_det:
                        frames = 0
                        if pipe is not None:
                            if not pipe.can_submit():          # (only reached when nothing is waiting: cannot happen, kept as a guard)
                                break
                            pipe.submit(frame, want_dets=False)
                    frames += 1
                    waiting.append((frame, is_det))
                    if pipe is None:
                        break                                  # no look-ahead without the fused pipeline
                if not waiting and not drawing:
                    break
                # ---- oldest frame: collect its result (detection frames), hand it to a drawing thread ----
                if waiting:
                    frame, is_det = waiting.popleft()
                    actions = []
                    if is_det:
                        if pipe is not None:
                            detections, _ = pipe.collect(want_dets=False)
                        else:
                            detections = self._track(frame)
                        if detections is not None and self.tracker is not None and self.action_id is not None:
                            actions = self.action_id.update(detections)
                        hold_detections = detections
                    drawing.append((pool.submit(self._draw, frame, hold_detections), hold_detections, actions))
                # ---- yield finished frames in order (keep up to draw_workers frames in the drawing threads) ----
                while drawing and (drawing[0][0].done() or len(drawing) > self.draw_workers or not waiting):
                    fut, hold, actions = drawing.popleft()
                    result = finish(fut.result())
                    yield result, hold, actions
                    if real_show and cv2.waitKey(1) & 0xFF == ord('q'):
                        eof = True
                        waiting.clear()
                        break
        finally:
            for p_ in self._pipelines.values():
                p_.drain()
            pool.shutdown(wait=True)
            reader.stop()
            vid.release()
            if out is not None:
                out.release()
            if real_show:
                cv2.destroyAllWindows()
