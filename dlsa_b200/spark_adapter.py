"""Spark adapter: the three calls of the reference's driver (projects/logistic_dlsa.py:293-345) with
the B200 engine underneath, for a Spark >= 3 job (``applyInPandas``; the reference's ``pandas_udf(..., GROUPED_MAP)`` form
of Spark 2.4 is the same function).  pyspark is imported lazily and only by the functions that need a SparkSession, so the
module loads (and its schema / UDF factory is testable) where Spark is absent.

    udf, schema = make_logistic_model_udf(usecols_x, Y_name, fit_intercept, dummy_info, dummy_factors_baseline, data_info)
    model_mapped_sdf = data_sdf.groupby("partition_id").applyInPandas(udf, schema=schema)      # :305-316 (executors with a GPU)
    Sig_inv_beta = dlsa.dlsa.dlsa_mapred(model_mapped_sdf)                                    # :337
    out_dlsa = dlsa.dlsa.dlsa(Sig_inv_beta.iloc[:, 2:], Sig_inv_beta["beta_byOLS"], n, fit_intercept)   # :341-344

``map_on_driver`` is the efficient form for one multi-GPU box: it pulls the frame through Arrow once and fits every
partition in one batched call instead of one Python worker call per group.
"""
from __future__ import annotations

from . import dummies, models


def output_fields(usecols_x, dummy_info, fit_intercept, dummy_factors_baseline):
    """Field list of the mapped frame: ``par_id`` (int), ``coef``, ``Sig_invMcoef`` and one double per design column —
    ``schema_beta`` of projects/logistic_dlsa.py:299-303 built on ``convert_schema`` (dlsa/utils_spark.py:5-33)."""
    return [("par_id", "int"), ("coef", "double"), ("Sig_invMcoef", "double")] + dummies.convert_schema(
        usecols_x, dummy_info, fit_intercept, dummy_factors_baseline)


def schema_ddl(fields) -> str:
    """The same schema as a DDL string (what ``applyInPandas(..., schema=...)`` accepts without importing pyspark types)."""
    return ", ".join(f"`{n}` {t}" for n, t in fields)


def make_logistic_model_udf(usecols_x, Y_name, fit_intercept=False, dummy_info=[], dummy_factors_baseline=[],
                            data_info=[]):
    """-> (function(pandas frame of one partition) -> p x (3+p) pandas frame, schema DDL string)."""
    fields = output_fields(usecols_x, dummy_info, fit_intercept, dummy_factors_baseline)
    cols = [n for n, _ in fields]

    def logistic_model_udf(sample_df):
        out = models.logistic_model(sample_df=sample_df, Y_name=Y_name, fit_intercept=fit_intercept, dummy_info=dummy_info,
                                    dummy_factors_baseline=dummy_factors_baseline, data_info=data_info)
        out = out[cols]                       # Spark matches applyInPandas output by position: pin the schema order
        out["par_id"] = out["par_id"].astype("int32")
        return out

    return logistic_model_udf, schema_ddl(fields)


def map_on_executors(data_sdf, usecols_x, Y_name, fit_intercept=False, dummy_info=[], dummy_factors_baseline=[],
                     data_info=[]):
    """``data_sdf.groupby("partition_id").applyInPandas(udf, schema)``: every executor's Python worker needs a visible
    B200.  Returns the lazy mapped Spark DataFrame (feed it to ``dlsa.dlsa.dlsa_mapred``)."""
    udf, schema = make_logistic_model_udf(usecols_x, Y_name, fit_intercept, dummy_info, dummy_factors_baseline, data_info)
    return data_sdf.groupby("partition_id").applyInPandas(udf, schema=schema)


def map_on_driver(data_sdf, Y_name, fit_intercept=False, dummy_info=[], dummy_factors_baseline=[], data_info=[]):
    """Collect through Arrow once (``toPandas()``; enable ``spark.sql.execution.arrow.pyspark.enabled``) and fit all
    partitions in one batched call on the driver's GPU: the stacked frames ``dlsa_mapred`` takes."""
    pdf = data_sdf.toPandas() if hasattr(data_sdf, "toPandas") else data_sdf
    return models.logistic_model_partitions(pdf, Y_name, fit_intercept, dummy_info, dummy_factors_baseline, data_info)
