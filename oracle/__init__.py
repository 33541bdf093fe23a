"""CPU oracle for the detect-and-track hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker (or as
the timed CPU baseline), never as the thing shipped.  The product path
(``yolo_deepsort_b200``) never imports this package and fails loudly when its
CUDA library is missing.

Each module restates one stage of GlassyWing/yolo_deepsort (reference paths are
relative to the reference checkout) with the same fp32 arithmetic, on CPU:

=====================  ==========================================================
module                 follows
=====================  ==========================================================
``darknet_ref``        yolo3/utils/parse_config.py:1-19, yolo3/models/models.py:25-102,
                       167-224, 292-366, yolo3/utils/model_build.py:12-19, 52-137,
                       317-332, yolo3/detect/img_detect.py:61-95
``nms_ref``            torchvision.ops.nms 0.26.0 (third party)
``cv_resize_ref``      cv2.resize INTER_LINEAR u8, OpenCV 4.13.0 (third party, App. C)
``reid_ref``           deep_sort/deep/model.py:5-95, deep_sort/deep/feature_extractor.py:12-58
``lsap_ref``/``lsap.c``  scipy.optimize.linear_sum_assignment 1.18.1 (third party, App. B)
``sort_ref``           deep_sort/sort/{kalman_filter,nn_matching,iou_matching,
                       linear_assignment,tracker,track,detection}.py, deep_sort/deep_sort.py:46-146
=====================  ==========================================================

Pinning: the reference ships no tests and one known-answer snippet
(deep_sort/sort/kalman_filter.py:259-273).  The oracle is therefore pinned by
(a) that snippet, (b) golden vectors produced by importing the *unmodified*
reference from /root/reference in the build container
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``, committed together with the
script), and (c) live differential fuzzing of the third-party restatements
against scipy / torchvision / cv2, which are present in the image
(``tests/test_oracle_*.py``).
"""
