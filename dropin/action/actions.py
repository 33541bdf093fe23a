from yolo_deepsort_b200.action import Action, BreakInto, FastCrossing, Glide, Landing, TakeOff  # noqa: F401  (action/actions.py)
