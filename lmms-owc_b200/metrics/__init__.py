"""Metric plugin API of the scoring path (mirror of the reference's ``src.data.metrics``)."""

from ._api import (AGGREGATIONS, METRICS, AggregationInfo, MetricInfo, get_aggregation, get_aggregation_builder,
                   get_aggregation_info, get_aggregations_info, get_metric, get_metric_builder, get_metric_info,
                   get_metrics_info, install_into_reference, register_aggregation, register_metric)
from ._group import GROUP_METRICS, semantic_matching  # noqa: F401  (registers the aggregations)
from ._instance import INSTANCE_METRICS  # noqa: F401  (registers the metrics; must follow _group)

__all__ = ["AGGREGATIONS", "METRICS", "GROUP_METRICS", "INSTANCE_METRICS", "AggregationInfo", "MetricInfo",
           "get_aggregation", "get_aggregation_builder", "get_aggregation_info", "get_aggregations_info",
           "get_metric", "get_metric_builder", "get_metric_info", "get_metrics_info", "install_into_reference",
           "register_aggregation", "register_metric", "semantic_matching"]
