"""Mirror of demos/yolov3_huaweiShip/utils (hot-path utilities only)."""
