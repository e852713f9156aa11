/*
 * plangpu_desc.h -- the flat plan descriptor passed to pg_plan_compile().
 *
 * The Go shim walks the PhysicalOperator subtree it wants to off-load
 * (/root/reference/pkg/compute/builder_physical_operator.go:49-66: Typ, Outputs,
 * Filters, Children, Info) and its Expr trees (pkg/compute/expr.go:49-60) and
 * appends int64 words.  No pointers, no Go structs.
 *
 *   desc  := PG_DESC_MAGIC PG_DESC_VERSION node
 *   node  := PG_OP_TOPK   nkeys (out_idx desc)* limit node      -- Limit <- Order [<- Project] <- Agg fused:
 *                         ORDER BY output columns of the aggregate below (executor_order.go,
 *                         sort_encoder.go:65-81 key semantics), keep the first `limit` rows
 *                         (executor_limit.go:120-137); limit < 0 = no limit.  Root only.
 *          | PG_OP_SCAN   slot nfilters expr*
 *          | PG_OP_FILTER nfilters expr* node
 *          | PG_OP_JOIN   jointype nconds (expr_probe expr_build)* nout (side idx)* node_probe node_build
 *          | PG_OP_PROJECT nexprs expr* node                 -- row-emitting pipelines only (root)
          | PG_OP_AGG    ngroups expr* naggs (aggfn ltype width scale expr)* nhaving expr* nout (kind idx)* node
 *                         -- (ltype width scale) = the aggregate's RESULT type, e.g. sum(DECIMAL(w,s)) ->
 *                            DECIMAL(38,s) (function_aggr.go:48-55); COUNT(*) passes ntokens = 0
 *   expr  := ntokens token*                       -- postfix (children before function)
 *   token := PG_TK_COL   side idx ltype width scale
 *          | PG_TK_CONST ltype width scale v0      -- INTEGER/BIGINT/DATE: value; DECIMAL: unscaled
 *                                                    value at `scale`; FLOAT/DOUBLE: IEEE-754 double bits
 *          | PG_TK_STR   nbytes word*              -- VARCHAR literal, bytes packed little-endian
 *          | PG_TK_FUNC  fn nargs ltype width scale
 *
 * Column references: in a SCAN's filters, side=0 and idx = column index in the
 * bound table.  Above an operator, side = child number (JOIN: 0 probe/left =
 * Children[0], 1 build/right = Children[1], executor_join.go:237-264) and idx =
 * position in that child's output list: SCAN outputs = all table columns; JOIN
 * outputs = its (side idx) list; AGG outputs = its (kind idx) list.
 * In an AGG's HAVING and output list: kind 0 = group key i, kind 1 = aggregate i
 * (expr_exec.go:248-265 maps ColRef.table()<0 to child chunks, >=0 to agg states).
 */
#ifndef PLANGPU_DESC_H
#define PLANGPU_DESC_H

#define PG_DESC_MAGIC 0x31504750 /* "PGP1" */
#define PG_DESC_VERSION 1

/* operators (POT_* of builder_physical_operator.go) */
#define PG_OP_SCAN 1
#define PG_OP_FILTER 2
#define PG_OP_JOIN 3
#define PG_OP_AGG 4
#define PG_OP_TOPK 5
#define PG_OP_PROJECT 6 /* POT_Project (executor_project.go:24-82): nexprs expr* node -- one output column per expression over the child's
                           outputs.  Root of a ROW-EMITTING pipeline: [PROJECT] <- [FILTER]* <- (SCAN | JOIN(scan, scan)) returns rows, not groups */

/* join types (LOT_JoinType) */
#define PG_JOIN_INNER 1
#define PG_JOIN_SEMI 2
#define PG_JOIN_ANTI 3
#define PG_JOIN_MARK 4
#define PG_JOIN_LEFT 5
#define PG_JOIN_ANTI_MARK 6 /* NOT EXISTS: the same mark column as MARK, filtered with mark = false (builder_plan.go:421-429) */

/* tokens */
#define PG_TK_COL 1
#define PG_TK_CONST 2
#define PG_TK_STR 3
#define PG_TK_FUNC 4

/* logical types (common.LTID_*, pkg/common/ltype.go) */
#define PG_LT_BOOLEAN 1
#define PG_LT_INTEGER 2
#define PG_LT_BIGINT 3
#define PG_LT_DATE 4
#define PG_LT_DECIMAL 5
#define PG_LT_FLOAT 6
#define PG_LT_DOUBLE 7
#define PG_LT_VARCHAR 8
#define PG_LT_HUGEINT 9

/* scalar functions (names in pkg/compute/function.go:89-128) */
#define PG_FN_ADD 1
#define PG_FN_SUB 2
#define PG_FN_MUL 3
#define PG_FN_DIV 4
#define PG_FN_EQ 10
#define PG_FN_NE 11
#define PG_FN_LT 12
#define PG_FN_LE 13
#define PG_FN_GT 14
#define PG_FN_GE 15
#define PG_FN_IN 16 /* in(arg, const...) (function_operator_boolean.go:393-504: INT32, VARCHAR) */
#define PG_FN_LIKE 17     /* like(arg, pattern): wildcardMatch, % = any run, _ = any byte (function_operator_boolean.go:336-381) */
#define PG_FN_NOT_LIKE 18
#define PG_FN_EXTRACT 19  /* extract('year', date) -> INTEGER (binStringInt32ExtractOp, function_operator_binary.go:259-265) */
#define PG_FN_AND 20
#define PG_FN_OR 21
#define PG_FN_NOT 22
#define PG_FN_CASE 23 /* case: args = [ELSE, WHEN1, THEN1, WHEN2, THEN2 ...] exactly Expr.Children of the reference's CASE
                         (executeCase, expr_exec.go:144-246); a missing ELSE is a PG_TK_CONST with ltype 0 (NULL) */
#define PG_FN_CAST 30 /* cast(arg AS ltype width scale) (function_cast.go:474-512) */

/* aggregate functions (function_aggr.go:26-165) */
#define PG_AGG_SUM 1
#define PG_AGG_AVG 2
#define PG_AGG_COUNT 3
#define PG_AGG_MIN 4
#define PG_AGG_MAX 5

#endif
