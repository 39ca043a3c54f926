"""Mirror of demos/yolov3_u/ (hot-path utilities only)."""
