"""Instance-level pass-through builders of the three similarity metrics.

Reference: ``src/data/metrics/_instance.py:294-312, 404-422, 445-462`` - the per-document function
is the identity; all work happens in the paired aggregation so documents can be batched.
"""

from ._api import register_metric

__all__ = ["INSTANCE_METRICS", "concept_semantic_similarity", "mean_average_semantic_similarity",
           "semantic_similarity"]

INSTANCE_METRICS = ["concept_semantic_similarity", "mean_average_semantic_similarity", "semantic_similarity"]


@register_metric(group_fn_name="concept_semantic_similarity", higher_is_better=True,
                 output_types=["generate_until"], can_bootstrap=False)
def concept_semantic_similarity(items: list) -> list:
    """Pass-through; see the aggregation of the same name."""
    return items


@register_metric(group_fn_name="mean_average_semantic_similarity", higher_is_better=True,
                 output_types=["generate_until"], can_bootstrap=False)
def mean_average_semantic_similarity(items: list) -> list:
    """Pass-through; see the aggregation of the same name."""
    return items


@register_metric(group_fn_name="semantic_similarity", higher_is_better=True, output_types=["generate_until"],
                 can_bootstrap=False)
def semantic_similarity(items: list) -> list:
    """Pass-through; see the aggregation of the same name."""
    return items
