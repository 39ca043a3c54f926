"""Mirror of demos/yolov3_u/utils (hot-path utilities only)."""
