from yolo_deepsort_b200.label_draw import LabelDrawer, draw_rects, draw_rects_and_labels, draw_single_img  # noqa: F401  (yolo3/utils/label_draw.py)
