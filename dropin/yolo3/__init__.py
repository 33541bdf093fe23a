"""Stands in for the reference package `yolo3` (only the modules video_deepsort.py touches)."""
