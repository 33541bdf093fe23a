from yolo_deepsort_b200.action import ActionIdentify  # noqa: F401  (action/action_Identify.py)
