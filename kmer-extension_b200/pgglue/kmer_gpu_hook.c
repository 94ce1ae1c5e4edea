/*
 * kmer_gpu_hook.c -- planner hook that routes the STOCK counting query to the GPU.
 *
 * The reference counts k-mers with plain SQL over its set-returning function (kmer--1.0.0.sql:101-104, plans at
 * kmer-tests.sql:1178-1181):
 *
 *   S1  SELECT k.kmer, count(*) FROM generate_kmers('ACGTACGT'::dna, 4) AS k(kmer) GROUP BY k.kmer;        -- TEST 13.1
 *   S2  SELECT kmer, count(*) FROM (SELECT generate_kmers(dna, 21) AS kmer FROM reads [WHERE ...]) s GROUP BY kmer;
 *   S3  SELECT k.kmer, count(*) FROM reads r, generate_kmers(r.dna, 21) AS k(kmer) GROUP BY k.kmer;
 *
 * i.e. one generate_kmers SRF scan per row (kmer.c:289-351) under a HashAggregate keyed through kmer_hash / kmer_equals
 * (kmer.c:226-245,353-365).  With this file in the module (and kmer_gpu.c's kmer_gpu_counts declared in SQL), the SAME query
 * text is planned as a Function Scan on
 *
 *       kmer_gpu_counts(ARRAY(SELECT dna FROM reads [WHERE ...]), 21)        -- S2, S3
 *       kmer_gpu_counts(ARRAY['ACGTACGT'::dna], 4)                           -- S1
 *
 * whose rows are exactly the groups of the original query: (kmer, count(*)::bigint), unordered like a HashAggregate's.  An
 * ORDER BY / LIMIT over the two output columns is kept (it only refers to target entries).  Everything else -- HAVING, extra
 * columns, other aggregates, count(DISTINCT ...), FILTER, a second SRF, WHERE on the outer level, WITH ORDINALITY, a parameter for k --
 * is left to PostgreSQL untouched.  `SET kmer.gpu_offload = off` disables the rewrite.  The counting query may itself be a
 * sub-select in FROM of a larger statement (offload_walk).
 *
 * Limit of this form: the column travels as ONE dna[] datum, and PostgreSQL caps a datum at 1 GB (MaxAllocSize); a table beyond
 * that needs the batched form (several kmer_gpu_counts calls merged, or a CustomScan feeding the library page by page) --
 * DESIGN.md section 8.
 *
 * The rewrite happens on the analyzed Query in planner_hook, before standard_planner(): no new plan node type, no executor
 * hook; the function scan runs kmer_gpu.c's SRF.  Semantics that make it exact: generate_kmers is STRICT (a NULL dna yields no
 * rows; kmer_gpu_counts skips NULL elements), IMMUTABLE and PARALLEL SAFE; a row shorter than k, k outside 1..32 or an invalid
 * base raise the reference's own errors through the library (include/kmer_cuda.h).
 *
 * Compiles against PostgreSQL >= 13 (planner_hook with query_string).  In this repository, which has no PostgreSQL, it is
 * compiled against oracle/pgshim and EXECUTED on hand-built Query trees by tests/c/hook_driver.c.
 */
#include "postgres.h"
#include "fmgr.h"
#include "catalog/namespace.h"
#include "catalog/pg_type.h"
#include "nodes/makefuncs.h"
#include "nodes/nodeFuncs.h"
#include "nodes/parsenodes.h"
#include "optimizer/planner.h"
#include "parser/parse_func.h"
#include "utils/guc.h"
#include "utils/lsyscache.h"

#define KMER_COUNT_STAR_OID 2803 /* pg_proc.dat: count(*) ; count("any") is 2147 and is not count(*) */

void _PG_init(void);
bool kmer_gpu_try_offload(Query *parse); /* also called by the test driver */

static bool kmer_gpu_offload = true;
static planner_hook_type prev_planner_hook = NULL;

typedef struct KmerCatalog
{
	Oid dna, dna_array, kmer;
	Oid generate_kmers; /* generate_kmers(dna, integer) */
	Oid gpu_counts;		/* kmer_gpu_counts(dna[], integer) */
} KmerCatalog;

/* Everything is looked up by name at plan time: the extension may be created, dropped or updated at any moment. */
static bool
kmer_catalog_lookup(KmerCatalog *c)
{
	Oid args[2];

	c->dna = TypenameGetTypid("dna");
	c->kmer = TypenameGetTypid("kmer");
	if (c->dna == InvalidOid || c->kmer == InvalidOid)
		return false;
	c->dna_array = get_array_type(c->dna);
	if (c->dna_array == InvalidOid)
		return false;
	args[0] = c->dna;
	args[1] = INT4OID;
	c->generate_kmers = LookupFuncName(list_make1(makeString("generate_kmers")), 2, args, true);
	args[0] = c->dna_array;
	c->gpu_counts = LookupFuncName(list_make1(makeString("kmer_gpu_counts")), 2, args, true);
	return c->generate_kmers != InvalidOid && c->gpu_counts != InvalidOid;
}

/* generate_kmers(<dna expression>, <non-null integer constant>) */
static bool
is_generate_kmers_call(Node *n, const KmerCatalog *c, Node **dna_arg, Const **k_arg)
{
	FuncExpr *f;
	Node *a0, *a1;

	if (!IsA(n, FuncExpr))
		return false;
	f = (FuncExpr *) n;
	if (f->funcid != c->generate_kmers || !f->funcretset || list_length(f->args) != 2)
		return false;
	a0 = (Node *) linitial(f->args);
	a1 = (Node *) lsecond(f->args);
	if (!IsA(a1, Const) || ((Const *) a1)->constisnull || ((Const *) a1)->consttype != INT4OID)
		return false;
	if (exprType(a0) != c->dna)
		return false;
	*dna_arg = a0;
	*k_arg = (Const *) a1;
	return true;
}

static bool
is_count_star(Node *n)
{
	Aggref *a;

	if (!IsA(n, Aggref))
		return false;
	a = (Aggref *) n;
	return a->aggfnoid == KMER_COUNT_STAR_OID && a->aggstar && a->args == NIL && a->aggdirectargs == NIL && a->aggorder == NIL &&
		   a->aggdistinct == NIL && a->aggfilter == NULL && a->agglevelsup == 0;
}

/* A plain SELECT level with nothing on it that the rewrite would have to carry along. */
static bool
is_plain_select(const Query *q)
{
	return q->commandType == CMD_SELECT && q->utilityStmt == NULL && q->resultRelation == 0 && q->cteList == NIL && !q->hasWindowFuncs &&
		   !q->hasDistinctOn && !q->hasRecursive && !q->hasModifyingCTE && !q->hasForUpdate && q->returningList == NIL &&
		   q->groupingSets == NIL && q->havingQual == NULL && q->windowClause == NIL && q->distinctClause == NIL && q->rowMarks == NIL &&
		   q->setOperations == NULL && q->jointree != NULL;
}

typedef struct KmerCountShape
{
	int shape;		 /* 1, 2, 3 as in the header comment */
	int kmer_varno;	 /* range table index of the k-mer column in the outer query */
	AttrNumber kmer_attno;
	Node *dna_arg;	 /* first argument of generate_kmers */
	Const *k_arg;	 /* second argument */
} KmerCountShape;

/*
 * The outer level: GROUP BY exactly the k-mer column, every target entry either that column or count(*).
 * On success the k-mer column is known as (varno, attno).
 */
static bool
outer_level_matches(const Query *q, int kmer_varno, AttrNumber kmer_attno, Oid kmer_type)
{
	ListCell *lc;
	Index group_ref;
	bool group_seen = false, count_seen = false;

	if (!is_plain_select(q) || !q->hasAggs || q->hasTargetSRFs || q->hasSubLinks || list_length(q->groupClause) != 1)
		return false;
	if (q->jointree->quals != NULL)
		return false;
	group_ref = ((SortGroupClause *) linitial(q->groupClause))->tleSortGroupRef;
	foreach (lc, q->targetList)
	{
		TargetEntry *tle = (TargetEntry *) lfirst(lc);
		Node *e = (Node *) tle->expr;

		if (is_count_star(e))
		{
			count_seen = true;
			continue;
		}
		if (IsA(e, Var))
		{
			Var *v = (Var *) e;

			if (v->varno == kmer_varno && v->varattno == kmer_attno && v->varlevelsup == 0 && v->vartype == kmer_type)
			{
				if (tle->ressortgroupref == group_ref)
					group_seen = true;
				continue;
			}
		}
		return false; /* anything else in the target list: not ours */
	}
	return group_seen && count_seen;
}

static bool
match_count_query(Query *q, const KmerCatalog *c, KmerCountShape *m)
{
	RangeTblEntry *rte1;

	if (q->jointree == NULL || q->rtable == NIL)
		return false;
	rte1 = (RangeTblEntry *) linitial(q->rtable);
	memset(m, 0, sizeof(*m));

	if (list_length(q->rtable) == 1 && list_length(q->jointree->fromlist) == 1 && IsA(linitial(q->jointree->fromlist), RangeTblRef) &&
		((RangeTblRef *) linitial(q->jointree->fromlist))->rtindex == 1)
	{
		if (rte1->rtekind == RTE_SUBQUERY && !rte1->lateral && rte1->subquery != NULL)
		{
			/* S2: the subquery's only output column is the SRF */
			Query *sq = rte1->subquery;
			TargetEntry *tle;

			if (!is_plain_select(sq) || sq->hasAggs || !sq->hasTargetSRFs || sq->groupClause != NIL || sq->sortClause != NIL ||
				sq->limitOffset != NULL || sq->limitCount != NULL || list_length(sq->targetList) != 1)
				return false;
			tle = (TargetEntry *) linitial(sq->targetList);
			if (tle->resjunk || !is_generate_kmers_call((Node *) tle->expr, c, &m->dna_arg, &m->k_arg))
				return false;
			m->shape = 2;
			m->kmer_varno = 1;
			m->kmer_attno = 1;
		}
		else if (rte1->rtekind == RTE_FUNCTION && !rte1->funcordinality && list_length(rte1->functions) == 1)
		{
			/* S1: a function scan over a constant dna */
			RangeTblFunction *rtf = (RangeTblFunction *) linitial(rte1->functions);

			if (!is_generate_kmers_call(rtf->funcexpr, c, &m->dna_arg, &m->k_arg) || !IsA(m->dna_arg, Const) ||
				((Const *) m->dna_arg)->constisnull)
				return false;
			m->shape = 1;
			m->kmer_varno = 1;
			m->kmer_attno = 1;
		}
		else
			return false;
	}
	else if (list_length(q->rtable) == 2 && list_length(q->jointree->fromlist) == 2 && IsA(linitial(q->jointree->fromlist), RangeTblRef) &&
			 IsA(lsecond(q->jointree->fromlist), RangeTblRef) && ((RangeTblRef *) linitial(q->jointree->fromlist))->rtindex == 1 &&
			 ((RangeTblRef *) lsecond(q->jointree->fromlist))->rtindex == 2)
	{
		/* S3: FROM <relation> r, generate_kmers(r.<dna column>, k) */
		RangeTblEntry *rte2 = (RangeTblEntry *) lsecond(q->rtable);
		RangeTblFunction *rtf;
		Var *v;

		if (rte1->rtekind != RTE_RELATION || rte2->rtekind != RTE_FUNCTION || rte2->funcordinality || list_length(rte2->functions) != 1)
			return false;
		rtf = (RangeTblFunction *) linitial(rte2->functions);
		if (!is_generate_kmers_call(rtf->funcexpr, c, &m->dna_arg, &m->k_arg) || !IsA(m->dna_arg, Var))
			return false;
		v = (Var *) m->dna_arg;
		if (v->varno != 1 || v->varlevelsup != 0 || v->varattno <= 0)
			return false;
		m->shape = 3;
		m->kmer_varno = 2;
		m->kmer_attno = 1;
	}
	else
		return false;
	return outer_level_matches(q, m->kmer_varno, m->kmer_attno, c->kmer);
}

/* FROM kmer_gpu_counts(<dna[] expression>, k) AS kmer_gpu_counts(kmer, count) */
static RangeTblEntry *
make_gpu_counts_rte(const KmerCatalog *c, Node *dna_array_expr, Const *k_arg)
{
	RangeTblEntry *rte = makeNode(RangeTblEntry);
	RangeTblFunction *rtf = makeNode(RangeTblFunction);
	FuncExpr *call = makeFuncExpr(c->gpu_counts, RECORDOID, list_make2(dna_array_expr, k_arg), InvalidOid, InvalidOid, COERCE_EXPLICIT_CALL);

	call->funcretset = true;
	rtf->funcexpr = (Node *) call;
	rtf->funccolcount = 2; /* OUT kmer kmer, OUT count bigint: the column types come from the function's catalog entry */
	rte->rtekind = RTE_FUNCTION;
	rte->functions = list_make1(rtf);
	rte->funcordinality = false;
	rte->eref = makeAlias("kmer_gpu_counts", list_make2(makeString("kmer"), makeString("count")));
	rte->lateral = false;
	rte->inh = false;
	rte->inFromCl = true;
	return rte;
}

static void
rewrite_count_query(Query *q, const KmerCatalog *c, const KmerCountShape *m)
{
	Node *dna_array_expr;
	ListCell *lc;
	bool group_ref_sorted = false;
	Index group_ref = ((SortGroupClause *) linitial(q->groupClause))->tleSortGroupRef;

	if (m->shape == 1)
	{
		ArrayExpr *arr = makeNode(ArrayExpr);

		arr->array_typeid = c->dna_array;
		arr->array_collid = InvalidOid;
		arr->element_typeid = c->dna;
		arr->elements = list_make1(m->dna_arg);
		arr->multidims = false;
		arr->location = -1;
		dna_array_expr = (Node *) arr;
	}
	else
	{
		Query *inner;
		SubLink *sub = makeNode(SubLink);

		if (m->shape == 2)
		{
			/* the subquery keeps its FROM / WHERE; its one output column becomes the dna expression itself */
			TargetEntry *tle;

			inner = ((RangeTblEntry *) linitial(q->rtable))->subquery;
			tle = (TargetEntry *) linitial(inner->targetList);
			tle->expr = (Expr *) m->dna_arg;
			tle->resname = "dna";
			inner->hasTargetSRFs = false;
		}
		else
		{
			/* SELECT r.<dna column> FROM <relation> r : the relation's range table entry moves one level down */
			RangeTblRef *ref = makeNode(RangeTblRef);

			inner = makeNode(Query);
			inner->commandType = CMD_SELECT;
			inner->querySource = QSRC_ORIGINAL;
			inner->canSetTag = true;
			inner->rtable = list_make1(linitial(q->rtable));
#if PG_VERSION_NUM >= 160000
			inner->rteperminfos = q->rteperminfos; /* the relation's permission info goes with it */
			q->rteperminfos = NIL;
#endif
			ref->rtindex = 1;
			inner->jointree = makeFromExpr(list_make1(ref), NULL);
			inner->targetList = list_make1(makeTargetEntry((Expr *) m->dna_arg, 1, "dna", false));
		}
		sub->subLinkType = ARRAY_SUBLINK;
		sub->subLinkId = 0;
		sub->testexpr = NULL;
		sub->operName = NIL;
		sub->subselect = (Node *) inner;
		sub->location = -1;
		dna_array_expr = (Node *) sub;
		q->hasSubLinks = true;
	}

	/* one range table entry: the function scan; both output columns are plain Vars of it */
	{
		RangeTblRef *ref = makeNode(RangeTblRef);

		ref->rtindex = 1;
		q->rtable = list_make1(make_gpu_counts_rte(c, dna_array_expr, m->k_arg));
		q->jointree = makeFromExpr(list_make1(ref), NULL);
	}
	foreach (lc, q->sortClause)
		if (((SortGroupClause *) lfirst(lc))->tleSortGroupRef == group_ref)
			group_ref_sorted = true;
	foreach (lc, q->targetList)
	{
		TargetEntry *tle = (TargetEntry *) lfirst(lc);

		if (is_count_star((Node *) tle->expr))
			tle->expr = (Expr *) makeVar(1, 2, INT8OID, -1, InvalidOid, 0);
		else
		{
			Var *v = (Var *) tle->expr;

			v->varno = 1;
			v->varattno = 1;
			if (tle->ressortgroupref == group_ref && !group_ref_sorted)
				tle->ressortgroupref = 0;
		}
		tle->resorigtbl = InvalidOid;
		tle->resorigcol = 0;
	}
	q->groupClause = NIL;
	q->hasAggs = false;
}

/*
 * The counting query need not be the statement itself: it is found at any depth of sub-selects in FROM, e.g.
 *   SELECT kmer::text, count FROM (SELECT kmer, count(*) AS count FROM (SELECT generate_kmers(dna, 21) ...) s GROUP BY kmer) t ORDER BY 2 DESC;
 * (CREATE TABLE AS / INSERT ... SELECT / EXPLAIN hand their SELECT to the planner as a statement of its own.)  A level that
 * matches is rewritten in place; its output columns keep their numbers, names and types, so the levels above do not notice.
 * Returns the number of levels rewritten.
 */
static int
offload_walk(Query *q, const KmerCatalog *cat, int depth)
{
	KmerCountShape m;
	ListCell *lc;
	int n = 0;

	if (q == NULL || depth > 16)
		return 0;
	if (q->commandType == CMD_SELECT && q->hasAggs && q->groupClause != NIL && match_count_query(q, cat, &m))
	{
		rewrite_count_query(q, cat, &m);
		return 1;
	}
	foreach (lc, q->rtable)
	{
		RangeTblEntry *rte = (RangeTblEntry *) lfirst(lc);

		if (rte->rtekind == RTE_SUBQUERY && rte->subquery != NULL)
			n += offload_walk(rte->subquery, cat, depth + 1);
	}
	return n;
}

/* does any level have an aggregate under a GROUP BY?  (cheap, before any catalog lookup) */
static bool
has_grouped_aggregate(const Query *q, int depth)
{
	ListCell *lc;

	if (q == NULL || depth > 16)
		return false;
	if (q->hasAggs && q->groupClause != NIL)
		return true;
	foreach (lc, q->rtable)
	{
		RangeTblEntry *rte = (RangeTblEntry *) lfirst(lc);

		if (rte->rtekind == RTE_SUBQUERY && has_grouped_aggregate(rte->subquery, depth + 1))
			return true;
	}
	return false;
}

/* Exposed for the test driver: true if some level of the query was rewritten. */
bool
kmer_gpu_try_offload(Query *parse)
{
	KmerCatalog cat;

	if (!kmer_gpu_offload || parse == NULL)
		return false;
	if (!has_grouped_aggregate(parse, 0))
		return false;
	if (!kmer_catalog_lookup(&cat))
		return false;
	return offload_walk(parse, &cat, 0) > 0;
}

static PlannedStmt *
kmer_gpu_planner(Query *parse, const char *query_string, int cursorOptions, ParamListInfo boundParams)
{
	(void) kmer_gpu_try_offload(parse);
	if (prev_planner_hook)
		return prev_planner_hook(parse, query_string, cursorOptions, boundParams);
	return standard_planner(parse, query_string, cursorOptions, boundParams);
}

void
_PG_init(void)
{
	DefineCustomBoolVariable("kmer.gpu_offload", "Plan generate_kmers(...) GROUP BY kmer / count(*) as a GPU function scan (kmer_gpu_counts).", NULL,
							 &kmer_gpu_offload, true, PGC_USERSET, 0, NULL, NULL, NULL);
	prev_planner_hook = planner_hook;
	planner_hook = kmer_gpu_planner;
}
