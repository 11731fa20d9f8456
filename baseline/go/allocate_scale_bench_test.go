// Scale benchmark of the allocate action against the reference's own harness. Written for this repository as the
// optional "reference-Go" timing of SURVEY.md §8(d); NOT compiled or run here (no Go toolchain in the image).
package allocate

import (
	"fmt"
	"testing"

	v1 "k8s.io/api/core/v1"

	schedulingv1 "volcano.sh/apis/pkg/apis/scheduling/v1beta1"
	"volcano.sh/volcano/cmd/scheduler/app/options"
	"volcano.sh/volcano/pkg/scheduler/api"
	"volcano.sh/volcano/pkg/scheduler/conf"
	"volcano.sh/volcano/pkg/scheduler/framework"
	"volcano.sh/volcano/pkg/scheduler/plugins/binpack"
	"volcano.sh/volcano/pkg/scheduler/plugins/gang"
	"volcano.sh/volcano/pkg/scheduler/plugins/nodeorder"
	"volcano.sh/volcano/pkg/scheduler/plugins/predicates"
	"volcano.sh/volcano/pkg/scheduler/plugins/priority"
	"volcano.sh/volcano/pkg/scheduler/uthelper"
	"volcano.sh/volcano/pkg/scheduler/util"
)

type lcg struct{ s uint64 }

func (l *lcg) next() uint64 { l.s = l.s*6364136223846793005 + 1442695040888963407; return l.s >> 33 }
func (l *lcg) pick(weights []int) int {
	total := 0
	for _, w := range weights {
		total += w
	}
	r := int(l.next() % uint64(total))
	for i, w := range weights {
		if r < w {
			return i
		}
		r -= w
	}
	return len(weights) - 1
}

func BenchmarkAllocateScale(b *testing.B) {
	const numNodes, numPods = 10000, 100000
	options.ServerOpts = options.NewServerOption()
	options.ServerOpts.PercentageOfNodesToFind = 100 // parity mode of the CUDA path
	options.ServerOpts.MinNodesToFind = 100
	options.ServerOpts.MinPercentageOfNodesToFind = 5

	rnd := &lcg{s: 20260921}
	skuCPU := []string{"32", "64", "96", "128"}
	skuMem := []string{"128Gi", "256Gi", "768Gi", "1024Gi"}
	skuPods := []string{"110", "110", "110", "250"}
	nodes := make([]*v1.Node, 0, numNodes)
	for i := 0; i < numNodes; i++ {
		k := rnd.pick([]int{40, 30, 20, 10})
		scalars := []api.ScalarResource{{Name: "pods", Value: skuPods[k]}}
		if k >= 2 {
			scalars = append(scalars, api.ScalarResource{Name: "nvidia.com/gpu", Value: "8"})
		}
		labels := map[string]string{"zone": fmt.Sprintf("z%d", rnd.next()%16), "pool": fmt.Sprintf("p%d", rnd.next()%8)}
		nodes = append(nodes, util.BuildNode(fmt.Sprintf("node-%06d", i), api.BuildResourceList(skuCPU[k], skuMem[k], scalars...), labels))
	}

	reqCPU := []string{"500m", "1", "2", "4", "8", "32"}
	reqMem := []string{"1Gi", "2Gi", "8Gi", "16Gi", "64Gi", "256Gi"}
	reqGPU := []string{"", "", "", "", "1", "8"}
	gangSizes := []int{1, 2, 4, 8, 16, 64, 256}
	gangWeights := []int{420, 210, 140, 105, 84, 70, 60} // ~ 1/rank
	var podGroups []*schedulingv1.PodGroup
	pods := make([]*v1.Pod, 0, numPods)
	for j := 0; len(pods) < numPods; j++ {
		size := gangSizes[rnd.pick(gangWeights)]
		if size > numPods-len(pods) {
			size = numPods - len(pods)
		}
		minMember := size
		if rnd.next()%10 >= 7 && size > 1 {
			minMember = size / 2
		}
		t := rnd.pick([]int{30, 30, 20, 10, 8, 2})
		pg := fmt.Sprintf("pg%d", j)
		podGroups = append(podGroups, util.BuildPodGroup(pg, "ns", "q1", int32(minMember), nil, schedulingv1.PodGroupInqueue))
		var scalars []api.ScalarResource
		if reqGPU[t] != "" {
			scalars = append(scalars, api.ScalarResource{Name: "nvidia.com/gpu", Value: reqGPU[t]})
		}
		selector := map[string]string{}
		if rnd.next()%4 == 0 {
			selector["zone"] = fmt.Sprintf("z%d", rnd.next()%16)
		}
		for k := 0; k < size; k++ {
			pods = append(pods, util.BuildPod("ns", fmt.Sprintf("%s-worker-%d", pg, k), "", v1.PodPending,
				api.BuildResourceList(reqCPU[t], reqMem[t], scalars...), pg, map[string]string{"volcano.sh/task-spec": "worker"}, selector))
		}
	}
	queues := []*schedulingv1.Queue{util.BuildQueue("q1", 1, nil)}

	plugins := map[string]framework.PluginBuilder{
		priority.PluginName: priority.New, gang.PluginName: gang.New, predicates.PluginName: predicates.New,
		nodeorder.PluginName: nodeorder.New, binpack.PluginName: binpack.New,
	}
	on := true
	tiers := []conf.Tier{
		{Plugins: []conf.PluginOption{
			{Name: priority.PluginName, EnabledJobOrder: &on, EnabledTaskOrder: &on},
			{Name: gang.PluginName, EnabledJobOrder: &on, EnabledJobReady: &on, EnabledJobPipelined: &on},
		}},
		{Plugins: []conf.PluginOption{
			{Name: predicates.PluginName, EnabledPredicate: &on},
			{Name: nodeorder.PluginName, EnabledNodeOrder: &on},
			{Name: binpack.PluginName, EnabledNodeOrder: &on, Arguments: map[string]interface{}{
				"binpack.weight": 10, "binpack.cpu": 5, "binpack.memory": 1,
				"binpack.resources": "nvidia.com/gpu", "binpack.resources.nvidia.com/gpu": 2}},
		}},
	}

	for i := 0; i < b.N; i++ {
		b.StopTimer()
		test := uthelper.TestCommonStruct{Name: "allocate-scale", Plugins: plugins, PodGroups: podGroups, Pods: pods, Nodes: nodes, Queues: queues}
		ssn := test.RegisterSession(tiers, nil)
		action := New()
		b.StartTimer()
		action.Execute(ssn) // the timed region: one allocate cycle
		b.StopTimer()
		test.Close()
	}
	b.ReportMetric(float64(numPods), "pending-pods/op")
}
