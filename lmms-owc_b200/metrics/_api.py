"""Metric plugin registry: same surface as the reference's ``src/data/metrics/_api.py``.

``register_metric`` / ``register_aggregation`` / ``get_metric_info`` and the ``MetricInfo`` /
``AggregationInfo`` records (reference: ``_api.py:101-109, 260-314``; ``src/schema/_base.py:18-136``)
are what ``eval_metrics.py:81-90`` and the task layer call; only the three similarity entries are
populated here. ``install_into_reference()`` swaps these implementations into an imported copy of
the reference's own registry (both dicts, because ``MetricInfo.group_fn`` is resolved at decoration
time - ``_api.py:303-310``).
"""

from __future__ import annotations

from collections.abc import Callable

from pydantic import BaseModel, Field

__all__ = ["AGGREGATIONS", "METRICS", "AggregationInfo", "MetricInfo", "register_aggregation", "register_metric",
           "get_aggregation", "get_aggregation_builder", "get_aggregation_info", "get_aggregations_info",
           "get_metric", "get_metric_builder", "get_metric_info", "get_metrics_info", "install_into_reference"]


class AggregationInfo(BaseModel):
    """Info about an available aggregation metric."""

    name: str = Field(description="Aggregation metric key.")
    builder_fn: Callable | None = Field(description="Function to build the aggregation metric.", exclude=True,
                                        default=None)
    can_bootstrap: bool = Field(description="Whether the standard error can be bootstrapped.", default=False)


class MetricInfo(BaseModel):
    """Info about an available metric."""

    name: str = Field(description="Metric key.")
    higher_is_better: bool | None = Field(description="Whether higher values are better.", default=None)
    builder_fn: Callable | None = Field(description="Function to build the metric.", exclude=True, default=None)
    group_fn: Callable | None = Field(description="Aggregation function used by the metric.", exclude=True,
                                      default=None)
    output_types: list[str] = Field(description="The output types of the metric.", default_factory=list)
    can_bootstrap: bool = Field(description="Whether the standard error can be bootstrapped.", default=False)


AGGREGATIONS: dict[str, AggregationInfo] = {}
METRICS: dict[str, MetricInfo] = {}


def register_aggregation(name: str | None = None, can_bootstrap: bool = False) -> Callable:
    """Register an aggregation function under ``name`` (default: the function's lower-cased name)."""

    def decorator(aggregation: Callable) -> Callable:
        key = name or aggregation.__name__.lower()
        AGGREGATIONS[key] = AggregationInfo(name=key, builder_fn=aggregation, can_bootstrap=can_bootstrap)
        return aggregation

    return decorator


def register_metric(name: str | None = None, group_fn_name: str | None = None, higher_is_better: bool | None = None,
                    output_types: list | None = None, can_bootstrap: bool = False) -> Callable:
    """Register a metric; its ``group_fn`` is looked up in ``AGGREGATIONS`` now, like the reference."""

    def decorator(metric: Callable) -> Callable:
        key = name or metric.__name__.lower()
        METRICS[key] = MetricInfo(name=key, higher_is_better=higher_is_better, builder_fn=metric,
                                  group_fn=AGGREGATIONS[group_fn_name].builder_fn,
                                  output_types=output_types or [], can_bootstrap=can_bootstrap)
        return metric

    return decorator


def get_aggregation(aggregation_id: str, **aggregation_kwargs):
    return AGGREGATIONS[aggregation_id].builder_fn(**aggregation_kwargs)


def get_aggregation_builder(aggregation_id: str) -> Callable:
    return AGGREGATIONS[aggregation_id].builder_fn


def get_aggregation_info(aggregation_id: str) -> AggregationInfo:
    return AGGREGATIONS[aggregation_id]


def get_aggregations_info() -> list[AggregationInfo]:
    return list(AGGREGATIONS.values())


def get_metric(metric_id: str, **metric_kwargs):
    return METRICS[metric_id].builder_fn(**metric_kwargs)


def get_metric_builder(metric_id: str) -> Callable:
    return METRICS[metric_id].builder_fn


def get_metric_info(metric_id: str) -> MetricInfo:
    return METRICS[metric_id]


def get_metrics_info() -> list[MetricInfo]:
    return list(METRICS.values())


def install_into_reference(ref_metrics_module) -> list[str]:
    """Patch an imported ``src.data.metrics`` (the reference's package) to use this build's functions.

    Replaces ``AGGREGATIONS[name].builder_fn`` and ``METRICS[name].group_fn`` for every metric this
    package registers; returns the patched names. See INTEGRATION.md.
    """
    patched = []
    for key, info in AGGREGATIONS.items():
        if key in ref_metrics_module.AGGREGATIONS:
            ref_metrics_module.AGGREGATIONS[key].builder_fn = info.builder_fn
            if key in ref_metrics_module.METRICS:
                ref_metrics_module.METRICS[key].group_fn = info.builder_fn
            patched.append(key)
    return patched
