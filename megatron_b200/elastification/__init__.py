from .elastic import ElasticBudget, ElasticController, BudgetSampler, distillation_step, extract_mlp_subnetwork  # noqa: F401
