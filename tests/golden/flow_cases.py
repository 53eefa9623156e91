"""The inference-flow parity cases shared by make_golden_flows.py (reference side) and tests/test_gpu_flows.py (CUDA side)."""
import numpy as np


def make_image(seed, h, w):
    """A BGR uint8 'micrograph' (smooth noise): only its bytes matter — they seed the fake predictor."""
    rng = np.random.default_rng(seed)
    import cv2
    base = rng.integers(0, 255, (h // 8 + 1, w // 8 + 1, 3), dtype=np.uint8)
    return cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)


P0 = dict(base_seed=1, n=26)
P1 = dict(base_seed=2, n=22)
CASES = {
    # run_class_specific_inference(predictor, image, target_class, small_classes, confidence_threshold, iou_threshold, settings, mode)
    "class_specific_large": dict(fn="run_class_specific_inference", image_seed=11, shape=(200, 260), predictors=[P0],
                                 args=(0, {1}), kwargs=dict(confidence_threshold=0.2, iou_threshold=0.7)),
    "class_specific_small": dict(fn="run_class_specific_inference", image_seed=12, shape=(180, 240), predictors=[P0],
                                 args=(1, {1}), kwargs=dict(confidence_threshold=0.3, iou_threshold=0.7,
                                                            class_specific_settings={"class_1": {"min_size": 3}})),
    "class_specific_input_scaled": dict(fn="run_class_specific_inference", image_seed=13, shape=(210, 250),
                                        predictors=[dict(base_seed=3, n=24, input_scale=0.8)], args=(0, set()),
                                        kwargs=dict(confidence_threshold=0.1, iou_threshold=0.6)),
    "class_specific_no_parallel": dict(fn="run_class_specific_inference", image_seed=14, shape=(160, 200), predictors=[P1], parallel=False,
                                       args=(0, {1}), kwargs=dict(confidence_threshold=0.2)),
    "class_specific_zero_score": dict(fn="run_class_specific_inference", image_seed=15, shape=(160, 200),
                                      predictors=[dict(base_seed=4, n=20, zero_score=True)], args=(0, set()),
                                      kwargs=dict(confidence_threshold=0.0)),
    "class_specific_none_left": dict(fn="run_class_specific_inference", image_seed=16, shape=(160, 200), predictors=[P1],
                                     args=(0, {1}), kwargs=dict(confidence_threshold=1.5)),
    "ensemble": dict(fn="run_ensemble_inference", image_seed=21, shape=(200, 260), predictors=[P0, P1], ensemble=True,
                     args=(0, {1}, 0.2, 0.5), kwargs={}),
    "ensemble_small": dict(fn="run_ensemble_inference", image_seed=22, shape=(190, 230), predictors=[P0, P1], ensemble=True,
                           args=(1, {1}, 0.25, 0.4), kwargs={}),
    "ensemble_empty": dict(fn="run_ensemble_inference", image_seed=23, shape=(150, 170), predictors=[P0, P1], ensemble=True,
                           args=(0, {1}, 1.5, 0.5), kwargs={}),
    "iterative": dict(fn="run_iterative_class_inference", image_seed=31, shape=(200, 260), predictors=[dict(base_seed=5, n=30)],
                      args=(0, {1}), kwargs=dict(confidence_threshold=0.2)),
    "iterative_small_minsize": dict(fn="run_iterative_class_inference", image_seed=32, shape=(180, 220), predictors=[P1],
                                    args=(1, {1}), kwargs=dict(confidence_threshold=0.2, min_crys_size=150)),
    "single_scale_1p5": dict(fn="process_single_scale", image_seed=41, shape=(160, 210), predictors=[P0],
                             args=(0, {1}, 0.2, 1.5), kwargs={}),
    "single_scale_0p7": dict(fn="process_single_scale", image_seed=42, shape=(190, 230), predictors=[P0],
                             args=(1, {1}, 0.2, 0.7), kwargs={}),
    "adaptive_multiscale": dict(fn="run_adaptive_multiscale_inference", image_seed=51, shape=(150, 190), predictors=[P0],
                                args=(0,), kwargs=dict(confidence_threshold=0.2, small_classes={1})),
    "tile_pipeline": dict(fn="tile_based_inference_pipeline", image_seed=61, shape=(200, 260), predictors=[P0],
                          args=(0, {1}, 0.2), kwargs=dict(tile_size=96, overlap_ratio=0.25, upscale_factor=2.0, iou_threshold=0.7)),
    "tile_pipeline_no_edge_filter": dict(fn="tile_based_inference_pipeline", image_seed=62, shape=(170, 230), predictors=[P1],
                                         args=(1, {1}, 0.25), kwargs=dict(tile_size=80, overlap_ratio=0.2, upscale_factor=1.5,
                                                                          edge_filter_enabled=False)),
    "tile_pipeline_ensemble": dict(fn="tile_based_inference_pipeline", image_seed=63, shape=(160, 200), predictors=[P0, P1], ensemble=True,
                                   args=(0, {1}, 0.2), kwargs=dict(tile_size=96, overlap_ratio=0.25, upscale_factor=2.0, iou_threshold=0.5)),
}
