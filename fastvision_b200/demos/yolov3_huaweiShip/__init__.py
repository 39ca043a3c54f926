"""Mirror of demos/yolov3_huaweiShip/ (hot-path utilities only)."""
